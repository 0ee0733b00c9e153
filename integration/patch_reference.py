#!/usr/bin/env python3
"""Splices the GPU seam (integration/bmbs_seam.h) into a SCRATCH COPY of the reference's Schema.cpp.

    python3 integration/patch_reference.py /tmp/scratch/Schema.cpp

The copy then calls libbmbs_gpu.so once per sub-block of reads (bmbs_map_batch_se / bmbs_map_batch_pe, include/bmbs.h) where the
stock workers seed and verify read by read; everything else -- Process_Reads batching, CIGAR, MAPQ, SAM text, the output queue
-- is the reference's own code, untouched.  Patched workers: Map_Single_Seq_split (single end, Schema.cpp:26763) and
Map_Pair_Seq_split_fast (--pe, :21460).  Regions are found by their text, not by line number; the script fails loudly when an
anchor is missing.  Nothing of the reference is stored in this repository: the script holds only the new lines.
oracle/build_ref.sh runs it and links the result twice: against libbmbs_gpu.so (oracle/_ref/bitmapperBS_gpu, GPU tests) and
against the CPU oracle behind the same C ABI (oracle/_ref/bitmapperBS_seam_cpu, so the splice itself is tested without a GPU).
"""
import sys

SE_SEEDING = r'''
			/* ---- bmbs seam: this sub-block was seeded, located, voted and verified in one library call (seam.map above) */
			error_threshold1 = thread_e_f * read_batch[i].length;
			if (error_threshold1 >= max_error_threshold) error_threshold1 = max_error_threshold;
			candidate_length = 0; get_error = -1; extra_seed_flag = 1; map_among_references = 0; second_best_diff = 0;
			{
				const bmbs_read_result& R = seam.res[i];
				is_mutiple_map = R.is_multiple_map;
				if (R.state == BMBS_EXACT_UNIQUE)
				{
					sprintf(cigar, "%dM", read_batch[i].length);
					output_sam_end_to_end_output_buffer(read_batch[i].name, read_batch[i].seq, read_batch[i].rseq, read_batch[i].qual,
						R.site, read_batch[i].length - 1, 0, 0, cigar, read_batch[i].length, &current_sub_buffer, &bam_buffer, &map_among_references, 0, 42);
					if (map_among_references == 0) { unique_matched_read++; matched_read++; total_bases = total_bases + read_batch[i].length; }
					i++;
					continue;
				}
				if (R.state == BMBS_MULTI_EXACT)
				{
					matched_read++;
					if (ambiguous_out == 1)
					{
						/* the interval of the whole read, for the reference's own walk over its rows */
						C_to_T_forward(read_batch[i].seq, bsSeq, read_batch[i].length, &C_site);
						number_of_hits = count_backward_as_much_1_terminate(bsSeq, read_batch[i].length, &top, &bot, &pre_top, &pre_bot, &match_length);
						locates = candidates;
						output_ambiguous_exact_map_output_buffer(read_batch[i].length, read_batch[i].seq, read_batch[i].rseq, read_batch[i].qual, cigar,
							read_batch[i].name, locates, top, number_of_hits, max_seed_matches, &match_length, &current_sub_buffer, &bam_buffer,
							&map_among_references, thread_id);
						if (map_among_references != 0) matched_read--;
					}
					i++;
					continue;
				}
				if (R.state == BMBS_ONE_MISMATCH) { candidates[0] = R.site; candidate_length = 1; extra_seed_flag = 0; one_mismatch_site = R.one_mismatch_pos; }
				else if (R.state == BMBS_VERIFY) candidate_length = R.n_cand;
			}
			/* ---- end of the seam */

'''

SE_VERIFY = r'''
				/* ---- bmbs seam: the windows come back site-sorted with votes, end_site and err; what is left is the reference's
				   vote sort and the reduction in that order */
				candidates_votes_length = seam.votes_of(i, candidates_votes);
				std::sort(candidates_votes, candidates_votes + candidates_votes_length, compare_seed_votes);
				bmbs_seam_reduce(candidates_votes, candidates_votes_length, is_mutiple_map != 0, &min_err, &min_err_index, &second_best_diff);
				/* ---- end of the seam */

'''

PE_BODY = r'''
			/* ---- bmbs seam: both mates seeded, pair-filtered and verified in the library call above */
			{
				const bmbs_read_result& R1 = seam.res[2 * i];
				const bmbs_read_result& R2 = seam.res[2 * i + 1];
				if (R1.n_cand == 0 || R2.n_cand == 0) { i++; continue; }
				candidates_votes_length1 = seam.votes_of(2 * i, candidates_votes1);
				candidates_votes_length2 = seam.votes_of(2 * i + 1, candidates_votes2);
				const bool res1 = bmbs_seam_resolved(R1), res2 = bmbs_seam_resolved(R2);
				if (res1 && res2) { best_mapp_occ1 = candidates_votes_length1; best_mapp_occ2 = candidates_votes_length2; }
				else if (!res1 && !res2)
				{
					if (candidates_votes_length1 <= candidates_votes_length2)
					{
						best_mapp_occ1 = bmbs_seam_keep_hits(candidates_votes1, candidates_votes_length1, error_threshold1);
						if (best_mapp_occ1 == 0) { i++; continue; }
						filter_pairs_single_side(best_mapp_occ1, candidates_votes_length2, &candidates_votes1, &candidates_votes2,
							&candidates_votes_length2, inner_maxDistance_pair, inner_minDistance_pair);
						best_mapp_occ2 = bmbs_seam_keep_hits(candidates_votes2, candidates_votes_length2, error_threshold2);
					}
					else
					{
						best_mapp_occ2 = bmbs_seam_keep_hits(candidates_votes2, candidates_votes_length2, error_threshold2);
						if (best_mapp_occ2 == 0) { i++; continue; }
						filter_pairs_single_side(best_mapp_occ2, candidates_votes_length1, &candidates_votes2, &candidates_votes1,
							&candidates_votes_length1, inner_maxDistance_pair, inner_minDistance_pair);
						best_mapp_occ1 = bmbs_seam_keep_hits(candidates_votes1, candidates_votes_length1, error_threshold1);
					}
				}
				else if (res1) { best_mapp_occ1 = candidates_votes_length1; best_mapp_occ2 = bmbs_seam_keep_hits(candidates_votes2, candidates_votes_length2, error_threshold2); }
				else { best_mapp_occ2 = candidates_votes_length2; best_mapp_occ1 = bmbs_seam_keep_hits(candidates_votes1, candidates_votes_length1, error_threshold1); }
			}
			/* ---- end of the seam */

'''


def find(s, what, start, end):
    p = s.find(what, start, end)
    if p < 0:
        raise SystemExit(f"patch_reference: anchor not found: {what!r}")
    return p


def line_start(s, p):
    return s.rfind("\n", 0, p) + 1


def main():
    path = sys.argv[1]
    s = open(path, errors="surrogateescape").read()
    # -------- the header, after the reference's own includes
    inc = find(s, '#include "bam_prase.h"', 0, len(s))
    eol = s.find("\n", inc) + 1
    s = s[:eol] + '#include "bmbs_seam.h"   /* integration/bmbs_seam.h: the GPU seam */\n' + s[eol:]

    # -------- single end: Map_Single_Seq_split (up to the pbat twin that follows it)
    f0 = find(s, "void* Map_Single_Seq_split(void* arg)", 0, len(s))
    f1 = find(s, "void* Map_Single_Seq_split_pbat(void* arg)", f0, len(s))
    # (4) verification + reduction: from the candidate sort to the line that starts the post-processing
    b0 = line_start(s, find(s, "std::sort(candidates, candidates + candidate_length);", f0, f1))
    b1 = line_start(s, find(s, "min_candidates_votes_length = 0;", b0, f1))
    s = s[:b0] + SE_VERIFY + s[b1:]
    f1 = find(s, "void* Map_Single_Seq_split_pbat(void* arg)", f0, len(s))
    # (3) seeding: from C_to_T_forward to the one-mismatch decision
    a0 = line_start(s, find(s, "C_to_T_forward(read_batch[i].seq, bsSeq, read_batch[i].length, &C_site);", f0, f1))
    a1 = line_start(s, find(s, "if (extra_seed_flag == 0", a0, f1))
    s = s[:a0] + SE_SEEDING + s[a1:]
    # (2) one library call per sub-block
    c = find(s, "read_batch = curr_sub_block.read;", f0, f1)
    eol = s.find("\n", c) + 1
    s = s[:eol] + "\t\tseam.map(read_batch, NULL, obtain_reads_num, 0);   /* bmbs seam: the whole sub-block in one call */\n" + s[eol:]
    # (1) the worker's block object
    c = find(s, "Read_buffer_single_sub_block curr_sub_block;", f0, f1)
    eol = s.find("\n", c) + 1
    s = s[:eol] + "\tbmbs_seam_block seam;   /* bmbs seam */\n" + s[eol:]

    # -------- paired end, fast mode: Map_Pair_Seq_split_fast (up to the sensitive worker that follows it)
    f0 = find(s, "void* Map_Pair_Seq_split_fast(void* arg)", 0, len(s))
    f1 = find(s, "void* Map_Pair_Seq_split(void* arg)", f0, len(s))
    loop = find(s, "file_flag = get_pe_reads_mul_thread(&curr_sub_block);", f0, f1)
    a0 = line_start(s, find(s, "get_candidates_muti_thread(", loop, f1))
    a1 = line_start(s, find(s, "mapping_pair = 0;", a0, f1))
    s = s[:a0] + PE_BODY + s[a1:]
    f1 = find(s, "void* Map_Pair_Seq_split(void* arg)", f0, len(s))
    c = find(s, "read_batch2 = curr_sub_block.read2;", loop, f1)
    eol = s.find("\n", c) + 1
    s = s[:eol] + "\t\tseam.map(read_batch1, read_batch2, obtain_reads_num, 0);   /* bmbs seam: the whole sub-block in one call */\n" + s[eol:]
    c = find(s, "Read_buffer_pe_sub_block curr_sub_block;", f0, loop)
    eol = s.find("\n", c) + 1
    s = s[:eol] + "\tbmbs_seam_block seam;   /* bmbs seam */\n" + s[eol:]
    open(path, "w", errors="surrogateescape").write(s)
    print("patch_reference: Map_Single_Seq_split and Map_Pair_Seq_split_fast now call bmbs_map_batch_se / _pe per sub-block")


if __name__ == "__main__":
    main()

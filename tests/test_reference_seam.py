"""The drop-in seam, proven with the reference's own host code: oracle/build_ref.sh splices integration/bmbs_seam.h into a scratch
copy of the reference's Schema.cpp (integration/patch_reference.py), so that Map_Single_Seq_split and Map_Pair_Seq_split_fast
call bmbs_map_batch_se / bmbs_map_batch_pe once per sub-block of reads; Process_Reads batching, CIGAR, MAPQ, SAM text and the
output queue stay the reference's.  The result must write the SAM the stock reference writes.

  not gpu   bitmapperBS_seam_cpu -- the C ABI served by the CPU oracle (oracle/seam_cpu_shim.cpp): tests the splice itself
  gpu       bitmapperBS_gpu      -- the same patched sources linked against libbmbs_gpu.so
"""
import subprocess

import pytest

from conftest import ROOT, sam_body

SETS = [("se100", ["--seq", "se100.fq"]), ("se250", ["--seq", "se250.fq"]),
        ("pe150", ["--seq1", "pe150_1.fq", "--seq2", "pe150_2.fq", "--pe"]),
        ("pe100h", ["--seq1", "pe100h_1.fq", "--seq2", "pe100h_2.fq", "--pe"])]
EXTRA = [("se100_flags", ["--seq", "se100.fq", "--unmapped_out", "--ambiguous_out"]),
         ("pe100h_flags", ["--seq1", "pe100h_1.fq", "--seq2", "pe100h_2.fq", "--pe", "--unmapped_out", "--ambiguous_out", "-e", "0.06", "--min", "50", "--max", "450"]),
         ("se250_e", ["--seq", "se250.fq", "-e", "0.04", "--seed", "25"])]


def run_sets(exe, golden, threads):
    for name, args in SETS:
        subprocess.run([str(exe), "--search", "genome.fa", *args, "-t", str(threads), "-o", "seam.sam", "--mapstats", "seam.stats"],
                       cwd=golden, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert sorted(sam_body(golden / "seam.sam")) == sorted(sam_body(golden / f"{name}.sam")), name
        assert (golden / "seam.stats").read_text() == (golden / f"ref_{name}.stats").read_text(), name


def run_against_live_reference(exe, ref, golden):
    for name, args in EXTRA:
        subprocess.run([str(ref), "--search", "genome.fa", *args, "-t", "1", "-o", "stock.sam", "--mapstats", "stock.stats"],
                       cwd=golden, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([str(exe), "--search", "genome.fa", *args, "-t", "3", "-o", "seam.sam", "--mapstats", "seam.stats"],
                       cwd=golden, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert sorted(sam_body(golden / "seam.sam")) == sorted(sam_body(golden / "stock.sam")), name
        assert (golden / "seam.stats").read_text() == (golden / "stock.stats").read_text(), name


def test_patched_reference_over_the_cpu_oracle_writes_the_stock_sam(golden, built):
    exe = ROOT / "oracle/_ref/bitmapperBS_seam_cpu"
    if not exe.exists():
        pytest.skip("oracle/_ref/bitmapperBS_seam_cpu absent (built by oracle/build_ref.sh where /root/reference exists)")
    run_sets(exe, golden, 1)
    run_sets(exe, golden, 4)
    if built["ref"].exists():
        run_against_live_reference(exe, built["ref"], golden)


@pytest.mark.gpu
def test_patched_reference_over_the_gpu_library_writes_the_stock_sam(golden, built):
    exe = ROOT / "oracle/_ref/bitmapperBS_gpu"
    if not exe.exists():
        pytest.skip("oracle/_ref/bitmapperBS_gpu absent (built by oracle/build_ref.sh where /root/reference exists)")
    run_sets(exe, golden, 1)
    run_sets(exe, golden, 4)
    if built["ref"].exists():
        run_against_live_reference(exe, built["ref"], golden)

/* oracle/ (test infrastructure): link-time stand-ins for the 16 htslib entry
 * points the reference's --bam path references.  The oracle never runs --bam
 * (SAM text is what parity is checked on), so building the
 * reference's vendored htslib (its own Makefile) is unnecessary; any call
 * lands in abort(). */
#include <stdio.h>
#include <stdlib.h>
#define STUB(name) void name(void) { fprintf(stderr, "oracle build: htslib stub %s called (--bam unsupported)\n", #name); abort(); }
STUB(bam_destroy1) STUB(bam_hdr_destroy) STUB(bam_init1) STUB(chhy_bam_write1_pure)
STUB(chhy_lazy_flush_pure) STUB(chhy_sam_hdr_read) STUB(finish_bam_output_buffer)
STUB(hts_close) STUB(hts_open) STUB(hts_open_format) STUB(init_multiple_buffer)
STUB(pop_buffer_bam) STUB(sam_hdr_read) STUB(sam_hdr_write) STUB(sam_parse1) STUB(sam_write1)

#!/bin/bash
# Diagnostic variants of libhpxfft_b200.so (never shipped; selected at run time with HPXFFT_B200_LIB=<path>):
#   nomath     : FP64 butterflies/twiddles removed  -> memory + shared-memory + barrier skeleton alone
#   wrap       : all global traffic aliased onto an L2-resident window -> arithmetic + shared memory alone
#   contig     : C=2 row CTAs store contiguous (wrong) bins -> cost of the interleaved 16-byte stores
#   noprefetch : without the L2 prefetch of the next long row
#   cw32       : 32-column tiles (512-byte segments) instead of 16
set -e
cd "$(dirname "$0")/../hpx-fft_b200"
python build.py -DHPXFFT_B200_DIAG_NOMATH --out=$PWD/libdiag_nomath.so &
python build.py -DHPXFFT_B200_DIAG_WRAP --out=$PWD/libdiag_wrap.so &
wait
python build.py -DHPXFFT_B200_DIAG_CONTIG_STORE --out=$PWD/libdiag_contig.so &
python build.py -DHPXFFT_B200_NO_ROW_PREFETCH --out=$PWD/libdiag_noprefetch.so &
wait
python build.py -DHPXFFT_B200_CW=32 --out=$PWD/libdiag_cw32.so
ls -la *.so

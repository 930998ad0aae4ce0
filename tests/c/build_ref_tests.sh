#!/bin/bash
# Compile the reference's own, UNMODIFIED test programs (read where they lie under
# /root/reference/tests) against pfft_b200's headers and library.  Only binaries are
# produced (oracle/_ref/bin/, git-ignored, shipped to the GPU box by gpurun); no reference
# source is copied into the repository.  Programs that fail to compile are listed and skipped.
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REF:-/root/reference}"
OUT="$ROOT/oracle/_ref/bin"
mkdir -p "$OUT"
: > "$OUT/BUILD_LOG"
LIST2D="minimal_check_c2c minimal_check_c2c_transposed simple_check_c2c simple_check_c2c_4d simple_check_c2c_4d_transposed
 simple_check_c2c_transposed simple_check_c2c_float simple_check_c2c_inplace simple_check_c2c_transposed_inplace
 simple_check_c2c_newarray simple_check_c2c_transposed_newarray
 simple_check_ousam_c2c simple_check_ousam_c2c_4d simple_check_ousam_c2c_4d_transposed simple_check_ousam_c2c_transposed
 simple_check_ousam_c2c_4d_newarray simple_check_ousam_c2c_4d_transposed_newarray
 simple_check_ousam_c2r simple_check_ousam_r2c simple_check_ousam_r2c_4d simple_check_ousam_r2c_4d_transposed
 simple_check_ousam_r2c_transposed simple_check_ousam_r2c_padded simple_check_ousam_r2c_4d_newarray
 simple_check_ousam_r2c_4d_transposed_newarray
 simple_check_r2c simple_check_r2c_4d simple_check_r2c_4d_transposed simple_check_r2c_transposed
 simple_check_r2c_newarray simple_check_r2c_transposed_newarray simple_check_r2c_padded_newarray
 simple_check_r2c_padded_transposed_newarray
 simple_check_c2r_c2c simple_check_c2r_c2c_shifted simple_check_c2r_c2c_ousam_shifted
 simple_check_r2r simple_check_r2r_4d simple_check_r2r_4d_transposed simple_check_r2r_transposed
 simple_check_ousam_r2r simple_check_ousam_r2r_transposed
 simple_check_ghost_c2c simple_check_ghost_r2c_input simple_check_ghost_r2c_input_padded simple_check_ghost_r2c_output
 time_c2c time_c2c_transposed"
LIST3D="simple_check_c2c_4d_on_3d simple_check_c2c_4d_on_3d_transposed simple_check_ousam_c2c_4d_on_3d
 simple_check_ousam_c2c_4d_on_3d_transposed simple_check_ousam_r2c_4d_on_3d simple_check_ousam_r2c_4d_on_3d_transposed
 simple_check_r2c_4d_on_3d simple_check_r2c_4d_on_3d_transposed simple_check_r2r_4d_on_3d simple_check_r2r_4d_on_3d_transposed
 simple_check_c2c_3d_on_3d simple_check_c2c_3d_on_3d_transposed simple_check_r2c_3d_on_3d simple_check_r2c_3d_on_3d_transposed
 simple_check_r2r_3d_on_3d simple_check_r2r_3d_on_3d_transposed simple_check_ghost_c2c_3d_on_3d"
: > "$OUT/LIST_2D"; : > "$OUT/LIST_3D"
for name in $LIST2D $LIST3D; do
  src="$REF/tests/$name.c"
  [ -f "$src" ] || continue
  if gcc -std=gnu99 -O1 -w -I"$ROOT/include" "$src" -o "$OUT/$name" \
       -L"$ROOT/pfft_b200/lib" -lpfft_b200 -lm -Wl,-rpath,'$ORIGIN/../../../pfft_b200/lib' >>"$OUT/BUILD_LOG" 2>&1; then
    if echo " $(echo $LIST3D) " | grep -q " $name "; then echo "$name" >> "$OUT/LIST_3D"; else echo "$name" >> "$OUT/LIST_2D"; fi
  else
    echo "SKIPPED (does not compile): $name" >> "$OUT/BUILD_LOG"
  fi
done
# the reference's benchmark program (tests/bench_c2c.c): command-line driven, run by its own test
: > "$OUT/LIST_BENCH"
if gcc -std=gnu99 -O1 -w -I"$ROOT/include" "$REF/tests/bench_c2c.c" -o "$OUT/bench_c2c" \
     -L"$ROOT/pfft_b200/lib" -lpfft_b200 -lm -Wl,-rpath,'$ORIGIN/../../../pfft_b200/lib' >>"$OUT/BUILD_LOG" 2>&1; then
  echo "bench_c2c" >> "$OUT/LIST_BENCH"
else
  echo "SKIPPED (does not compile): bench_c2c" >> "$OUT/BUILD_LOG"
fi
echo "built $(cat "$OUT/LIST_2D" "$OUT/LIST_3D" "$OUT/LIST_BENCH" | wc -l) reference test programs into $OUT"

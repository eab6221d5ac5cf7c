#!/usr/bin/env bash
# ORACLE / TEST INFRASTRUCTURE.  Compiles the reference's own hot-path sources, unmodified and
# in place from /root/reference, against oracle/shim/ into oracle/_ref/libgrref.so.
# Outputs go ONLY to oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
# The reference's build system (cmake/autotools + Boost/FFTW/SWIG/Py2) is NOT run.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${GR_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/gnuradio-core" ]; then
  if [ -f "$OUT/libgrref.so" ]; then echo "build_ref: reference absent, keeping prebuilt $OUT/libgrref.so"; exit 0; fi
  echo "build_ref: $REF not found and no prebuilt library" >&2; exit 1
fi
mkdir -p "$OUT/gen" "$OUT/obj"
python3 "$HERE/expand_templates.py" "$REF" "$OUT/gen"
CORE="$REF/gnuradio-core/src/lib"
INC=(-I"$HERE/shim" -I"$OUT/gen" -I"$REF/gruel/src/include" -I"$CORE/general" -I"$CORE/runtime"
     -I"$CORE/filter" -I"$REF/gr-digital/include" -I"$REF/gr-pager/lib")
CXXFLAGS=(-O2 -fPIC -std=gnu++11 -w -include cmath -include cstdio -include cassert
          -Dgnuradio_core_EXPORTS -Dgnuradio_digital_EXPORTS -Dgnuradio_pager_EXPORTS)
SRCS=(
  "$OUT/gen/gr_fir_ccf.cc" "$OUT/gen/gr_fir_ccf_generic.cc"
  "$OUT/gen/gr_fir_fff.cc" "$OUT/gen/gr_fir_fff_generic.cc"
  "$OUT/gen/gr_fir_ccc.cc" "$OUT/gen/gr_fir_ccc_generic.cc"
  "$OUT/gen/gr_fir_filter_ccf.cc" "$OUT/gen/gr_fir_filter_fff.cc"
  "$OUT/gen/gr_freq_xlating_fir_filter_ccf.cc"
  "$CORE/filter/gr_fir_ccf_simd.cc" "$CORE/filter/gr_fir_ccf_x86.cc"
  "$CORE/filter/gr_fir_fff_simd.cc" "$CORE/filter/gr_fir_fff_x86.cc"
  "$CORE/filter/gr_fir_ccc_simd.cc" "$CORE/filter/gr_fir_ccc_x86.cc"
  "$CORE/filter/gr_pfb_channelizer_ccf.cc" "$CORE/filter/gr_pfb_arb_resampler_ccf.cc" "$CORE/filter/gr_pfb_decimator_ccf.cc"
  "$CORE/filter/gr_fft_filter_ccc.cc" "$CORE/filter/gri_fft_filter_ccc_generic.cc" "$CORE/general/gr_framer_sink_1.cc"
  "$CORE/filter/gri_mmse_fir_interpolator.cc"
  "$CORE/general/gr_reverse.cc" "$CORE/general/gr_fast_atan2f.cc" "$CORE/general/gr_count_bits.cc"
  "$CORE/general/gr_quadrature_demod_cf.cc" "$CORE/general/gr_fft_vcc.cc" "$CORE/general/gr_fft_vcc_fftw.cc"
  "$CORE/general/gr_firdes.cc" "$CORE/general/gr_remez.cc" "$CORE/general/gr_stream_to_streams.cc" "$CORE/general/gr_vector_to_streams.cc" "$CORE/general/gr_map_bb.cc" "$CORE/general/gr_unpack_k_bits_bb.cc"
  "$REF/gr-digital/lib/digital_clock_recovery_mm_ff.cc" "$REF/gr-digital/lib/digital_clock_recovery_mm_cc.cc"
  "$CORE/filter/gri_mmse_fir_interpolator_cc.cc"
  "$REF/gr-digital/lib/digital_correlate_access_code_bb.cc"
  "$REF/gr-digital/lib/digital_binary_slicer_fb.cc"
  "$REF/gr-pager/lib/pager_slicer_fb.cc"
  "$HERE/ref_capi.cc" "$HERE/ref_bench.cc"
)
OBJS=()
pids=()
for s in "${SRCS[@]}"; do
  o="$OUT/obj/$(basename "${s%.*}").o"; OBJS+=("$o")
  g++ "${CXXFLAGS[@]}" "${INC[@]}" -c "$s" -o "$o" &
  pids+=($!)
done
for s in "$CORE/general/malloc16.c" "$CORE/filter/gr_sincos.c"; do
  o="$OUT/obj/$(basename "${s%.*}").o"; OBJS+=("$o")
  gcc -O2 -fPIC -w -I"$CORE/general" -I"$CORE/filter" -I"$REF/gruel/src/include" -Dgnuradio_core_EXPORTS -c "$s" -o "$o" &
  pids+=($!)
done
for a in fcomplex_dotprod_sse64 float_dotprod_sse64 ccomplex_dotprod_sse64 \
         fcomplex_dotprod_3dnow64 float_dotprod_3dnow64 ccomplex_dotprod_3dnow64 ccomplex_dotprod_3dnowext64; do
  o="$OUT/obj/$a.o"; OBJS+=("$o")
  gcc -c -fPIC -I"$CORE/filter" "$CORE/filter/$a.S" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -o "$OUT/libgrref.so" "${OBJS[@]}" -lpthread
echo "build_ref: built $OUT/libgrref.so"

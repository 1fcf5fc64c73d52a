// Conv3d 3x3x3 / pad 1 (TMA-staged fp32 SIMT, conv_tma.cuh): "flat" full-row tiles for the volume widths of the
// LPBA / Mindboggle pyramids that are not a multiple of the 16/32-voxel warp tile (reference ModeT/models.py:186-228:
// every AvgPool3d(2) halves W, so 160 -> 80 -> 40 -> 20 and 224 -> 112 -> 56 -> 28).
#include <cstdio>
#include <cstdlib>

#include "conv_tma.cuh"

namespace smile {
using namespace conv_tma_detail;

namespace {
template <int FW, int NW>
int dispatch_flat(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                  double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps, cudaStream_t st) {
  constexpr int TH = (NW * 32) / FW;
  const long long plane_tiles = (long long)ceil_div(H, TH) * B;
  const long long want = 2LL * kNumSMs;
  auto ctas = [&](int V, int CO) { return plane_tiles * ceil_div(D, V) * ceil_div(Cout, CO); };
  // channel tile that divides Cout (measured with tools/tune_conv.py: 12 output channels on a 16-wide tile waste a
  // quarter of the FMAs -- 6->12 @80x96x80 114 us vs 79 us on 4-wide tiles)
  const int co_cap = Cout <= 4 ? 4 : (Cout <= 8 ? 8 : (Cout % 16 == 0 ? 16 : (Cout % 8 == 0 ? 8 : 4)));
#define SMILE_FCONV(CO, V, CIC) \
  return launch_tcfg<CO, V, 32, NW, CIC, FW>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st)
  if (const char* fc = getenv("SMILE_CONV_FORCE")) {  // tuning knob (tools/tune_conv.py): "CO:V" of the register tile
    int fco = 0, fv = 0;
    if (sscanf(fc, "%d:%d", &fco, &fv) == 2) {
      if (fco == 16 && fv == 4) SMILE_FCONV(16, 4, 4);
      if (fco == 16 && fv == 2) SMILE_FCONV(16, 2, 4);
      if (fco == 8 && fv == 4) SMILE_FCONV(8, 4, 4);
      if (fco == 8 && fv == 2) SMILE_FCONV(8, 2, 4);
      if (fco == 4 && fv == 8) SMILE_FCONV(4, 8, 2);
      if (fco == 4 && fv == 4) SMILE_FCONV(4, 4, 4);
      if (fco == 4 && fv == 2) SMILE_FCONV(4, 2, 4);
    }
  }
  // largest register tile (V depths x CO channels per thread) that still gives about two waves of CTAs; when
  // nothing does, the combination with the most CTAs
  if (co_cap == 16 && ctas(4, 16) >= want) SMILE_FCONV(16, 4, 4);
  if (co_cap >= 8 && ctas(4, 8) >= want) SMILE_FCONV(8, 4, 4);
  if (co_cap == 16 && ctas(2, 16) >= want) SMILE_FCONV(16, 2, 4);
  if (co_cap >= 8 && ctas(2, 8) >= want) SMILE_FCONV(8, 2, 4);
  if (co_cap == 4 && ctas(8, 4) >= want) SMILE_FCONV(4, 8, 2);
  if (co_cap == 4 && ctas(4, 4) >= want) SMILE_FCONV(4, 4, 4);
  if (co_cap >= 8 && ctas(2, 4) <= ctas(2, 8)) SMILE_FCONV(8, 2, 4);
  SMILE_FCONV(4, 2, 4);
#undef SMILE_FCONV
}
}  // namespace

// Full-row tiles when they waste fewer lanes than cutting W into 16/32-voxel tiles.
int launch_conv3d_tma_flat(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                           double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                           cudaStream_t st, bool* handled) {
  *handled = false;
  if (get_encode() == nullptr || (long long)Cin * D * H * W * 4 >= (1LL << 40)) return SMILE_OK;
  if (Cout <= 8 && W <= 40 && getenv("SMILE_CONV_FORCE") == nullptr) return SMILE_OK;  // measured: the tiled path is ~10 % ahead here
  *handled = true;
  switch (W) {
    case 20: return dispatch_flat<20, 4>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
    case 28: return dispatch_flat<28, 4>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
    case 40: return dispatch_flat<40, 4>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
    case 56: return dispatch_flat<56, 4>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
    case 80: return dispatch_flat<80, 8>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
    case 112: return dispatch_flat<112, 8>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
    default: break;
  }
  *handled = false;
  return SMILE_OK;
}

}  // namespace smile

// Conv3d 3x3x3 / pad 1, fp32 SIMT with TMA-staged input tiles: shape dispatch (kernel in conv_tma.cuh).
#include <cstdio>
#include <cstdlib>

#include "conv_tma.cuh"

namespace smile {
using namespace conv_tma_detail;

// Returns with *handled = false when the shape is not eligible (W % 4 != 0, tiny volumes); the
// caller then uses the generic kernel of conv.cu.
int launch_conv3d_tma(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                      double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                      cudaStream_t st, bool* handled) {
  *handled = false;
  if (W % 4 != 0 || W < 16 || get_encode() == nullptr) return SMILE_OK;
  if ((long long)Cin * D * H * W * 4 >= (1LL << 40)) return SMILE_OK;
  {
    bool flat = false;
    const bool no_flat = getenv("SMILE_CONV_NO_FLAT") != nullptr;  // tuning knob (tools/tune_conv.py), read per call
    if (!no_flat) {
      int rc = launch_conv3d_tma_flat(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st,
                                      &flat);
      if (flat) {
        *handled = true;
        return rc;
      }
    }
  }
  *handled = true;
#define SMILE_TCONV(CO, V, TWL, NW, CIC) \
  return launch_tcfg<CO, V, TWL, NW, CIC>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st)
  // Pick the largest register tile (V depths x CO channels per thread) that still yields about two
  // waves of CTAs; small volumes (coarse pyramid levels, CWM) would otherwise run on a handful of SMs.
  const int TWL = (W >= 64 || W == 32) ? 32 : 16;
  const int TH = 8;
  const long long plane_tiles = (long long)ceil_div(H, TH) * ceil_div(W, TWL) * B;
  const long long want = 2LL * kNumSMs;
  auto ctas = [&](int V, int CO) { return plane_tiles * ceil_div(D, V) * ceil_div(Cout, CO); };
  const int co_cap = Cout <= 4 ? 4 : (Cout <= 8 ? 8 : 16);
  // candidates in decreasing tile size: (V, CO)
  const int cand[6][2] = {{4, 16}, {8, 8}, {4, 8}, {2, 16}, {2, 8}, {2, 4}};
  int V = 2, CO = 4;
  if (co_cap == 4) {
    V = (ctas(8, 4) >= want) ? 8 : ((ctas(4, 4) >= want) ? 4 : 2);
    CO = 4;
  } else {
    bool found = false;
    for (int i = 0; i < 6 && !found; ++i) {
      if (cand[i][1] > co_cap) continue;
      if (ctas(cand[i][0], cand[i][1]) >= want) {
        V = cand[i][0];
        CO = cand[i][1];
        found = true;
      }
    }
    if (!found) {  // nothing reaches two waves: maximise the CTA count
      V = 2;
      CO = (co_cap >= 8 && ctas(2, 4) <= ctas(2, 8)) ? 8 : 4;
    }
  }
  if (const char* f = getenv("SMILE_CONV_FORCE")) {  // tuning knob: "CO:V" of the register tile
    int fco = 0, fv = 0;
    if (sscanf(f, "%d:%d", &fco, &fv) == 2 && fco <= co_cap * 2) {
      CO = fco;
      V = fv;
    }
  }
  if (TWL == 32) {
    if (CO == 4 && V == 8 && Cin == 1) SMILE_TCONV(4, 8, 32, 8, 1);  // first layer: do not stage / multiply a phantom channel
    if (CO == 4 && V == 8) SMILE_TCONV(4, 8, 32, 8, 2);
    if (CO == 4 && V == 4) SMILE_TCONV(4, 4, 32, 8, 4);
    if (CO == 4) SMILE_TCONV(4, 2, 32, 8, 4);
    if (CO == 8 && V == 8) SMILE_TCONV(8, 8, 32, 8, 2);
    if (CO == 8 && V == 4) SMILE_TCONV(8, 4, 32, 8, 4);
    if (CO == 8) SMILE_TCONV(8, 2, 32, 8, 4);
    if (V == 4) SMILE_TCONV(16, 4, 32, 8, 2);
    SMILE_TCONV(16, 2, 32, 8, 4);
  }
  if (CO == 4 && V == 8) SMILE_TCONV(4, 8, 16, 4, 2);
  if (CO == 4 && V == 4) SMILE_TCONV(4, 4, 16, 4, 4);
  if (CO == 4) SMILE_TCONV(4, 2, 16, 4, 4);
  if (CO == 8 && V == 8) SMILE_TCONV(8, 8, 16, 4, 2);
  if (CO == 8 && V == 4) SMILE_TCONV(8, 4, 16, 4, 4);
  if (CO == 8) SMILE_TCONV(8, 2, 16, 4, 4);
  if (V == 4) SMILE_TCONV(16, 4, 16, 4, 4);
  SMILE_TCONV(16, 2, 16, 4, 4);
#undef SMILE_TCONV
}

}  // namespace smile

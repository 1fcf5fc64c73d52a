/* TEST INFRASTRUCTURE (oracle/): checks that the 3-instruction division used by the fast kernels,
 *     q0 = p * rc;  r = fma(-q0, c, p);  q = fma(r, rc, q0)      with rc = RN(1 / c),
 * equals the IEEE-754 correctly rounded p / c that the reference computes in
 * SpatialTransformer.forward (ModeT/models.py:56: new_locs / (shape - 1)) for every volume extent c = S - 1
 * and a dense set of coordinates p (integers +- a few ulps, uniform randoms, wide exponent range).
 * Prints "tot=<n> bad=<m>"; exit status 0 iff bad == 0.   gcc -O2 -mfma -ffp-contract=off */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float mdiv(float p, float c, float rc) {
  float q0 = p * rc;
  float r = fmaf(-q0, c, p);
  return fmaf(r, rc, q0);
}
int main(int argc, char** argv) {
  int per = argc > 1 ? atoi(argv[1]) : 100000;
  uint64_t bad = 0, tot = 0, s = 88172645463325252ULL;
  for (int ci = 1; ci <= 1024; ci++) {
    volatile float c = (float)ci;
    float rc = 1.0f / c;
    for (int i = -20; i <= ci + 20; i++)
      for (int u = -3; u <= 3; u++) {
        float p = (float)i;
        uint32_t b;
        memcpy(&b, &p, 4);
        b += u;
        memcpy(&p, &b, 4);
        if (!isfinite(p)) continue;
        float a = p / c, m = mdiv(p, c, rc);
        tot++;
        if (memcmp(&a, &m, 4)) bad++;
      }
    for (int k = 0; k < per; k++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      float p = (float)(((double)(s >> 11) / 9007199254740992.0) * (ci + 40.0) - 20.0);
      float a = p / c, m = mdiv(p, c, rc);
      tot++;
      if (memcmp(&a, &m, 4)) bad++;
    }
    for (int k = 0; k < per / 4; k++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      uint32_t b = (uint32_t)s;
      b = (b & 0x807fffffu) | ((uint32_t)(64 + (s >> 40) % 128) << 23);
      float p;
      memcpy(&p, &b, 4);
      float a = p / c, m = mdiv(p, c, rc);
      tot++;
      if (memcmp(&a, &m, 4)) bad++;
    }
  }
  printf("tot=%llu bad=%llu\n", (unsigned long long)tot, (unsigned long long)bad);
  return bad != 0;
}

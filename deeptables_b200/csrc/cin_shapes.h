// Host-side shape bookkeeping for CIN (CIN.build, layers.py:638-680).
#pragma once
#include <cstddef>
#include <cstdint>

namespace dtb {

constexpr int kCinMaxLayers = 8;

struct CinShape {
  int F = 0, D = 0, n_layers = 0, direct = 0;
  int L[kCinMaxLayers];        // layer sizes
  int H[kCinMaxLayers + 1];    // field_nums: H[0] = F, H[k+1] = direct ? L[k] : L[k]/2
  int pool_lo[kCinMaxLayers];  // first feature map of layer k that goes to the pooled output
  int pool_n[kCinMaxLayers];   // how many
  int pcol0[kCinMaxLayers];    // its first column in pooled
  size_t w_off[kCinMaxLayers]; // float offset of filter k in the concatenated weights
  size_t b_off[kCinMaxLayers]; // float offset of bias k
  int P = 0, sumL = 0, Lmax = 0, Hmax = 0, Kmax = 0;
  size_t w_total = 0;

  // returns false when the reference would reject the configuration (layers.py:668-670)
  bool init(int F_, int D_, const int* sizes, int n, int direct_) {
    if (F_ <= 0 || D_ <= 0 || n <= 0 || n > kCinMaxLayers || !sizes) return false;
    F = F_; D = D_; n_layers = n; direct = direct_;
    H[0] = F;
    P = 0; sumL = 0; Lmax = 0; Hmax = F; Kmax = 0; w_total = 0;
    size_t boff = 0;
    for (int k = 0; k < n; ++k) {
      L[k] = sizes[k];
      if (L[k] <= 0) return false;
      const bool last = (k == n - 1);
      if (direct) {
        H[k + 1] = L[k];
        pool_lo[k] = 0; pool_n[k] = L[k];
      } else {
        if (!last && (L[k] % 2)) return false;
        H[k + 1] = L[k] / 2;
        pool_lo[k] = last ? 0 : L[k] / 2;
        pool_n[k] = last ? L[k] : L[k] / 2;
      }
      pcol0[k] = P;
      P += pool_n[k];
      w_off[k] = w_total;
      w_total += (size_t)F * H[k] * L[k];
      b_off[k] = boff;
      boff += L[k];
      sumL += L[k];
      if (L[k] > Lmax) Lmax = L[k];
      if (H[k] > Hmax) Hmax = H[k];
      if (F * H[k] > Kmax) Kmax = F * H[k];
    }
    return true;
  }
};

}  // namespace dtb

// Device-side restatement of Localizer:inputToFeatureRect + extract_roi_pooling_input's crop, shared by the detector
// and the training kernels.
#pragma once
#include "detect.h"

namespace frcnn {

// ------------------------------------------------------------------------------------------------ ROI geometry
// Localizer:inputToFeatureRect (Localizer.lua:41-67) including its dW/dH mix-ups, then the clip and the 1-based
// crop of extract_roi_pooling_input (objective.lua:5-13).  Returns false where the reference would raise.
__device__ __forceinline__ double lua_mod(double a, double b) { return a - floor(a / b) * b; }
// x / d for a small positive integer d: a multiplication by the exactly representable reciprocal when d is a power of
// two (bit-identical to the division), the division otherwise
__device__ __forceinline__ double div_int(double x, int d) {
  return (d & (d - 1)) == 0 ? x * (1.0 / (double)d) : x / (double)d;
}
__device__ __forceinline__ double lua_mod_int(double a, int b) { return a - floor(div_int(a, b)) * (double)b; }

__device__ inline bool roi_crop(const LocalizerDev& loc, double minX, double minY, double maxX, double maxY, int FH, int FW, int* y0,
                         int* y1, int* x0, int* x1) {
  for (int i = 0; i < loc.n; ++i) {
    const int ikW = loc.l[i][0], ikH = loc.l[i][1], idW = loc.l[i][2], idH = loc.l[i][3];
    const double kW = ikW, kH = ikH, dW = idW, dH = idH, pW = loc.l[i][4], pH = loc.l[i][5];
    if (dW < kW) {
      minX -= (kW - dW); maxX += (kW - dW);
      minY -= (kH - dH); maxY += (kH - dH);
    }
    minX += pW; maxX += pW;
    minY += pH; maxY += pH;
    minX = div_int(minX, idH);  // sic (Localizer.lua:52)
    minY = div_int(minY, idH);
    if (lua_mod_int(maxX - kW, idW) == 0.0) maxX = fmax(div_int(maxX - kW, idW) + 1.0, minX + 1.0);
    else maxX = fmax(ceil(div_int(maxX - kW, idW)) + 1.0, minX + 1.0);
    if (lua_mod_int(maxY - kH, idH) == 0.0) maxY = fmax(div_int(maxY - kH, idW) + 1.0, minY + 1.0);  // sic: / dW (Localizer.lua:60)
    else maxY = fmax(ceil(div_int(maxY - kH, idH)) + 1.0, minY + 1.0);
  }
  minX = floor(minX); minY = floor(minY); maxX = ceil(maxX); maxY = ceil(maxY);  // snapToInt (Rect.lua:147-149)
  // r:clip(Rect.new(0, 0, W, H)) (Rect.lua:73-80)
  const double cminX = fmin(fmax(minX, 0.0), (double)FW), cminY = fmin(fmax(minY, 0.0), (double)FH);
  const double cmaxX = fmax(fmin(maxX, (double)FW), 0.0), cmaxY = fmax(fmin(maxY, (double)FH), 0.0);
  // idx = { {}, { min(minY + 1, maxY), maxY }, { min(minX + 1, maxX), maxX } }  (1-based inclusive)
  const double ylo = fmin(cminY + 1.0, cmaxY), xlo = fmin(cminX + 1.0, cmaxX);
  if (ylo < 1.0 || xlo < 1.0 || cmaxY < ylo || cmaxX < xlo) return false;
  *y0 = (int)ylo - 1; *y1 = (int)cmaxY; *x0 = (int)xlo - 1; *x1 = (int)cmaxX;
  return true;
}


}  // namespace frcnn

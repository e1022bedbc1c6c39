#include "kgb_scorevalue.h"

#include <cmath>

namespace kgb {

std::vector<double> makeExpectedSVTable() {
  const double twoOverPi = 0.63661977236758134308;
  const int stepsPerUnit = 10, boundStdevs = 5;
  const int loStdev = -boundStdevs * stepsPerUnit, hiStdev = boundStdevs * stepsPerUnit;
  std::vector<double> pdf(hiStdev - loStdev + 1);
  for(int i = loStdev; i <= hiStdev; i++) {
    const double x = (double)i / stepsPerUnit;
    pdf[i - loStdev] = std::exp(-0.5 * x * x);
  }
  const int loSV = -(SV_MEAN_RADIUS * stepsPerUnit + stepsPerUnit / 2 + boundStdevs * SV_STDEV_LEN * stepsPerUnit), hiSV = -loSV;
  std::vector<double> sv(hiSV - loSV + 1);
  for(int i = loSV; i <= hiSV; i++) {
    const double score = (double)i / stepsPerUnit;
    sv[i - loSV] = std::atan((score - 0.0) / (1.0 * SV_ASSUMED_BSIZE)) * twoOverPi;
  }
  std::vector<double> table((size_t)SV_MEAN_LEN * SV_STDEV_LEN);
  for(int m = 0; m < SV_MEAN_LEN; m++) {
    const int meanSteps = (m - SV_MEAN_RADIUS) * stepsPerUnit - stepsPerUnit / 2;
    for(int sd = 0; sd < SV_STDEV_LEN; sd++) {
      double wSum = 0.0, wsvSum = 0.0;
      for(int i = loStdev; i <= hiStdev; i++) {
        const double w = pdf[i - loStdev];
        wSum += w;
        wsvSum += w * sv[meanSteps + sd * i - loSV];
      }
      table[(size_t)m * SV_STDEV_LEN + sd] = wsvSum / wSum;
    }
  }
  return table;
}

}  // namespace kgb

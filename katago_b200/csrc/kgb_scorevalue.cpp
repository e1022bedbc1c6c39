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

// ------------------------------------------------------------------------------------------------------------
// Student-t CDF table of the value weighting (search.cpp:131-137: DistributionTable over tdistcdf(z, 3), z in [-50, 50],
// 2000 points; core/fancymath.cpp:12-116).  tdistcdf goes through the regularized incomplete beta function, evaluated by
// the modified Lentz continued fraction with the textbook terms, tolerance 1e-15.
// ------------------------------------------------------------------------------------------------------------
namespace kgb {
namespace {

double lentzIncompleteBetaFraction(double x, double a, double b) {
  const double tiny = 1e-300, tolerance = 1e-15;
  double ret = 1.0, c = 1.0, d = 0.0;
  for(int n = 1; n < 100000; n++) {
    double num;
    if(n % 2 == 0) {
      const double m = n / 2;
      num = m * (b - m) * x / (a + 2.0 * m - 1.0) / (a + 2.0 * m);
    }
    else {
      const double m = (n - 1) / 2;
      num = -(a + m) * (a + b + m) * x / (a + 2.0 * m) / (a + 2.0 * m + 1.0);
    }
    d = 1.0 + num * d;
    if(d == 0.0) d = tiny;
    c = 1.0 + num / c;
    if(c == 0) c = tiny;
    d = 1.0 / d;
    const double mult = c * d;
    ret = ret * mult;
    if(std::fabs(mult - 1.0) <= tolerance) break;
  }
  return ret;
}

double regularizedIncompleteBeta(double x, double a, double b) {
  if(!(x >= 0.0 && x <= 1.0 && a > 0.0 && b > 0.0)) return NAN;
  if(x <= 0.0) return 0.0;
  if(x >= 1.0) return 1.0;
  const double logx = std::log(x), logy = std::log(1 - x);
  const double logbeta = std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b);
  if(x <= (a + 1.0) / (a + b + 2.0)) return std::exp(logx * a + logy * b - logbeta) / a / lentzIncompleteBetaFraction(x, a, b);
  return 1.0 - (std::exp(logy * b + logx * a - logbeta) / b / lentzIncompleteBetaFraction(1.0 - x, b, a));
}

double tdistcdf(double x, double v) {
  if(x >= 0) return 1.0 - regularizedIncompleteBeta(v / (x * x + v), v / 2.0, 0.5) / 2.0;
  return regularizedIncompleteBeta(v / (x * x + v), v / 2.0, 0.5) / 2.0;
}

}  // namespace

std::vector<double> makeValueWeightCdfTable() {
  std::vector<double> t(VW_TABLE_SIZE);
  for(int i = 0; i < VW_TABLE_SIZE; i++) {
    if(i == 0) t[i] = 0.0;
    else if(i == VW_TABLE_SIZE - 1) t[i] = 1.0;
    else {
      const double z = VW_MIN_Z + i * (VW_MAX_Z - VW_MIN_Z) / (double)(VW_TABLE_SIZE - 1);
      t[i] = tdistcdf(z, 3.0);
    }
  }
  return t;
}

}  // namespace kgb

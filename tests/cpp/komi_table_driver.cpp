// tests/cpp/komi_table_driver.cpp - TEST INFRASTRUCTURE for integration/b200_komi.h (tests/test_komi_search.py).
// stdin: "X Y N" then N lines "komi lead winLoss" (the reference's searches of every komi of a position, tests/golden/komitable.json.gz), then
// lines "startKomi".  For every start komi: computeLead driven like the C++ host drives it - the oracle knows nothing at first, every NeedKomi is
// answered from the table and the function is run again - printing "lead <%.9g> asked <komi> <komi> ...".
#include <cstdio>
#include <iostream>
#include <map>
#include <vector>

#include "integration/b200_komi.h"

int main() {
  int x, y, n;
  if(!(std::cin >> x >> y >> n)) return 1;
  std::map<float, std::pair<double, double>> table;
  for(int i = 0; i < n; i++) { double k, lead, wl; std::cin >> k >> lead >> wl; table[(float)k] = {lead, wl}; }
  double start;
  while(std::cin >> start) {
    b200::KomiOracle oracle(x, y);
    std::vector<float> asked;
    for(;;) {
      try {
        const float lead = b200::computeLead(start, oracle);
        std::printf("lead %.9g asked", (double)lead);
        for(float k : asked) std::printf(" %.1f", (double)k);
        std::printf("\n");
        break;
      }
      catch(const b200::NeedKomi& need) {
        auto it = table.find(need.komi);
        if(it == table.end()) { std::printf("error: komi %.1f is not in the table\n", (double)need.komi); return 1; }
        oracle.add(need.komi, it->second.first, it->second.second);
        asked.push_back(need.komi);
      }
    }
  }
  return 0;
}

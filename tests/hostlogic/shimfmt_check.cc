// TEST HARNESS (CPU): the C++ host shim's value formatting (plan_b200/host/chunk.hpp), fed from stdin:
//   D coef scale neg type_scale   -> decimal_value_string
//   F bits64                      -> go_float_string of the double with that bit pattern
//   T days                        -> date_string
// one answer per line.  tests/test_host_cpu.py compares the answers with the oracle's restatement.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include "../../plan_b200/host/chunk.hpp"

int main()
{
    std::string kind;
    while (std::cin >> kind) {
        if (kind == "D") {
            unsigned long long coef; int scale, neg, ts;
            std::cin >> coef >> scale >> neg >> ts;
            printf("%s\n", planhost::decimal_value_string((uint64_t)coef, scale, neg != 0, ts).c_str());
        } else if (kind == "F") {
            unsigned long long bits; double x;
            std::cin >> bits;
            memcpy(&x, &bits, 8);
            printf("%s\n", planhost::go_float_string(x).c_str());
        } else {
            int days;
            std::cin >> days;
            printf("%s\n", planhost::date_string(days).c_str());
        }
    }
    return 0;
}

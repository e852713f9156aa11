// Host-only harness over the PRODUCT's host-side arithmetic (plan_b200/csrc/hostdec.hpp, plan_ir.hpp): reads cases on
// stdin, prints results.  tests/test_host_cpu.py compares them with the oracle's independent restatements on the CPU.
//   Q <coef_a> <scale_a> <neg_a> <coef_b> <scale_b> <neg_b>   ->  "ok <coef> <scale> <neg>" | "fail"       (hd_quo)
//   N <neg> <hi> <lo> <scale>                                  ->  "ok <coef> <scale> <neg>" | "fail"       (hd_normalise of a u128)
//   W <pattern> <target>   ('~' stands for the empty string)   ->  "<match> <is_contains_fast_path>"        (wildcard_match)
//   F <op> <float literal bits as u32> <scale> <vmin> <vmax>    ->  "<lo> <hi>"   (float_cmp_to_range: cast(DECIMAL AS FLOAT) <op> literal)
#include <stdio.h>
#include <string.h>

#include <iostream>
#include <string>

#include "../../plan_b200/csrc/hostdec.hpp"
#include "../../plan_b200/csrc/plan_ir.hpp"

using namespace pg;

int main()
{
    std::string kind;
    while (std::cin >> kind) {
        if (kind == "Q") {
            unsigned long long ca, cb;
            int sa, na, sb, nb;
            std::cin >> ca >> sa >> na >> cb >> sb >> nb;
            HDec a, b, q;
            a.coef = ca; a.scale = sa; a.neg = na != 0;
            b.coef = cb; b.scale = sb; b.neg = nb != 0;
            if (hd_quo(a, b, &q)) printf("ok %llu %d %d\n", (unsigned long long)q.coef, q.scale, q.neg ? 1 : 0);
            else printf("fail\n");
        } else if (kind == "N") {
            int neg, scale;
            unsigned long long hi, lo;
            std::cin >> neg >> hi >> lo >> scale;
            HDec d;
            if (hd_normalise(neg != 0, ((u128)hi << 64) | (u128)lo, scale, &d)) printf("ok %llu %d %d\n", (unsigned long long)d.coef, d.scale, d.neg ? 1 : 0);
            else printf("fail\n");
        } else if (kind == "F") {
            int op, scale;
            unsigned bits;
            long long vmin, vmax;
            std::cin >> op >> bits >> scale >> vmin >> vmax;
            float k;
            memcpy(&k, &bits, 4);
            i64 lo, hi;
            float_cmp_to_range(op, k, scale, vmin, vmax, &lo, &hi);
            printf("%lld %lld\n", (long long)lo, (long long)hi);
        } else if (kind == "W") {
            std::string p, t, lit;
            std::cin >> p >> t;
            if (p == "~") p.clear();
            if (t == "~") t.clear();
            printf("%d %d\n", wildcard_match(p.data(), p.size(), t.data(), t.size()) ? 1 : 0, like_is_contains(p, &lit) ? 1 : 0);
        }
    }
    return 0;
}

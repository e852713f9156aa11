// Host-only, AddressSanitizer / UBSan build of the PRODUCT's descriptor parser (plan_b200/csrc/plan_ir.hpp DescReader): reads
// descriptors from stdin ("<nwords> w0 w1 ..." per line), parses each, prints "ok" or "bad".  tests/test_host_cpu.py feeds it
// thousands of truncated / mutated descriptors: any out-of-bounds read or undefined behaviour aborts the process.
#include <stdio.h>

#include <iostream>
#include <vector>

#include "../../plan_b200/csrc/plan_ir.hpp"

int main()
{
    long long n;
    while (std::cin >> n) {
        std::vector<int64_t> w((size_t)n);
        for (auto &x : w) { long long v; std::cin >> v; x = v; }
        bool ok = false;
        if (w.size() >= 3 && w[0] == PG_DESC_MAGIC && w[1] == PG_DESC_VERSION) {
            pg::DescReader rd(w.data() + 2, w.size() - 2);
            pg::Node root;
            ok = rd.node(&root) && rd.ok() && rd.pos() == w.size() - 2;
        }
        puts(ok ? "ok" : "bad");
    }
    return 0;
}

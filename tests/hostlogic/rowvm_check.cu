// Host-only harness over the PRODUCT's row-program compiler (plan_b200/csrc/rowvm_compile.hpp): no kernel is launched, no
// device memory touched.  stdin:
//   line 1: ncols, then per column "<pg_type> <scale> <ndict> <dict entries...>" (entries without blanks; '_' stands for a blank)
//   line 2: mode ("filters" | "exprs"), then the int64 words of a descriptor whose root is PG_OP_SCAN (mode filters: its filter
//           list is compiled as ONE conjunction) or PG_OP_PROJECT over a scan (mode exprs: every projection is one program)
// stdout: "ok" + the listing ("pre <col> <mask> <lo> <hi>" lines, "ins <pc> <op> <a> <b> <imm>" lines, "prog <p0> <p1> <kind>")
//         or "fail <reason>".  tests/test_host_cpu.py checks structure: which conjuncts became inline pre-tests, jump targets,
//         evaluation-stack depth on every path, refusals.
#include <stdio.h>

#include <iostream>
#include <string>
#include <vector>

#include "../../plan_b200/csrc/rowvm_compile.hpp"

using namespace pg;

int main()
{
    int ncols;
    std::cin >> ncols;
    pg_table t;
    t.name = "t";
    t.nrows = 1000;
    t.sealed = true;
    static char fake[64][16];                    // distinct non-null "device" addresses: the compiler only compares them
    for (int i = 0; i < ncols; i++) {
        Column c;
        int nd;
        std::cin >> c.type >> c.scale >> nd;
        c.name = "c" + std::to_string(i);
        c.width = 15;
        for (int k = 0; k < nd; k++) { std::string e; std::cin >> e; for (auto &ch : e) if (ch == '_') ch = ' '; c.dict.push_back(e); }
        c.d_data = fake[i];
        c.stats_ok = true;
        c.vmin = 0; c.vmax = 1000;
        for (int k = 0; k < 8; k++) c.present[k] = 0xffffffffu;
        t.cols.push_back(c);
    }
    std::string mode;
    std::cin >> mode;
    std::vector<int64_t> words;
    long long w;
    while (std::cin >> w) words.push_back(w);
    if (words.size() < 3 || words[0] != PG_DESC_MAGIC) { printf("fail bad descriptor\n"); return 0; }
    DescReader rd(words.data() + 2, words.size() - 2);
    Node root;
    if (!rd.node(&root) || !rd.ok()) { printf("fail malformed descriptor\n"); return 0; }
    RvCompiler cc;
    cc.tables[0] = &t;
    Resolver scope = [&](int idx, Src *s) { if (idx < 0 || idx >= (int)t.cols.size()) return false; s->side = 0; s->col = idx; s->mark = false; return true; };
    std::vector<std::string> progs;
    if (mode == "filters") {
        const Node *scan = &root;
        while (scan->op != PG_OP_SCAN && !scan->children.empty()) scan = &scan->children[0];
        std::vector<const Expr *> fl;
        for (auto &f : scan->filters) fl.push_back(&f);
        int p0, p1;
        if (!cc.compile_filters(fl, scope, &p0, &p1)) { printf("fail %s\n", cc.why.c_str()); return 0; }
        progs.push_back("prog " + std::to_string(p0) + " " + std::to_string(p1) + " " + std::to_string((int)RVK_BOOL));
    } else {
        if (root.op != PG_OP_PROJECT) { printf("fail root is not a projection\n"); return 0; }
        for (auto &e : root.exprs) {
            int k = 0;
            const int p0 = cc.ncode;
            if (!cc.compile(e, scope, &k)) { printf("fail %s\n", cc.why.c_str()); return 0; }
            progs.push_back("prog " + std::to_string(p0) + " " + std::to_string(cc.ncode) + " " + std::to_string(k) + " scale_bound " + std::to_string(cc.scale_bound(e, scope)));
        }
    }
    printf("ok\n");
    for (int i = 0; i < cc.npre; i++) {
        const RvPre &q = cc.code.pre[i];
        int col = -1;
        for (int c = 0; c < ncols; c++) if (cc.code.cols[q.col].col.p == (const void *)fake[c]) col = c;
        unsigned long long mask_lo = 0;
        if (q.mask >= 0) mask_lo = (unsigned long long)cc.code.masks[q.mask][0] | ((unsigned long long)cc.code.masks[q.mask][1] << 32);
        printf("pre %d %d %lld %lld %llu\n", col, q.mask, (long long)q.lo, (long long)q.hi, mask_lo);
    }
    for (int pc = 0; pc < cc.ncode; pc++) printf("ins %d %d %d %d %lld\n", pc, cc.code.ins[pc].op, cc.code.ins[pc].a, cc.code.ins[pc].b, (long long)cc.code.ins[pc].imm);
    for (auto &p : progs) printf("%s\n", p.c_str());
    return 0;
}

// Host-only harness over the PRODUCT's row-program compiler (plan_b200/csrc/rowvm_compile.hpp): no kernel is launched, no
// device memory touched.  stdin:
//   line 1: ncols, then per column "<pg_type> <scale> <ndict> <dict entries...>" (entries without blanks; '_' stands for a blank)
//   line 2: mode ("filters" | "exprs"), then the int64 words of a descriptor whose root is PG_OP_SCAN (mode filters: its filter
//           list is compiled as ONE conjunction) or PG_OP_PROJECT over a scan (mode exprs: every projection is one program)
//   mode "lower": the root is PG_OP_AGG over a scan; prints what the LOWERED kernels are given (plan_ir.hpp): "range <col> <lo> <hi>
//           <is_set> <like>" per filter column (lower_filters) and "agg <i> vscale <s> factors <col> <c> <s> <scale> ..." per aggregate
//           argument (lower_affprod), or "fail <reason>" when the shape does not lower (the expression kernels take it then)
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
    if (mode == "lower") {
        if (root.op != PG_OP_AGG) { printf("fail root is not an aggregate\n"); return 0; }
        const Node *scan = &root.children[0];
        while (scan->op != PG_OP_SCAN && !scan->children.empty()) scan = &scan->children[0];
        LowerCtx cx;
        cx.table = &t;
        cx.allow_nulls = true;
        std::vector<Range> ranges;
        if (!lower_filters(cx, scan->filters, ranges)) { printf("fail filters: %s\n", cx.why.c_str()); return 0; }
        std::vector<AffProd> args(root.aggs.size());
        for (size_t i = 0; i < root.aggs.size(); i++)
            if (!root.aggs[i].star && root.aggs[i].fn != PG_AGG_COUNT && !lower_affprod(cx, root.aggs[i].arg, args[i])) { printf("fail aggregate %zu: %s\n", i, cx.why.c_str()); return 0; }
        printf("ok\n");
        for (auto &r : ranges) printf("range %d %lld %lld %d %d\n", r.col, (long long)r.lo, (long long)r.hi, r.is_set ? 1 : 0, r.like);
        for (size_t i = 0; i < args.size(); i++) {
            printf("agg %zu vscale %d factors", i, args[i].vscale());
            for (auto &f : args[i].f) printf(" %d %lld %d %d", f.col, (long long)f.c, f.s, f.scale);
            printf("\n");
        }
        return 0;
    }
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

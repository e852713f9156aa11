// TEST HARNESS (CPU): the descriptor words the C++ host shim (plan_b200/host/gpu_exec.hpp + tpch_plans.hpp) serialises for its TPC-H
// plans, one line per query: "<name> <slot table names in slot order, comma separated> <words...>".  tests/test_host_cpu.py compares them
// with the Python mirror's serialisation of the same plans.  Links nothing: pg_last_error is stubbed (only PlanError uses it).
#include <cstdio>
#include <string>
#include <vector>
#include "../../plan_b200/host/chunk.hpp"
extern "C" const char *pg_last_error(void) { return ""; }
#include "../../plan_b200/host/gpu_exec.hpp"
#include "../../plan_b200/host/tpch_plans.hpp"

using namespace planhost;

static void dump(const char *name, const PhysicalOperator &op)
{
    Serializer s;
    s.plan(op);
    std::vector<std::string> byslot(s.slots.size());
    for (auto &kv : s.slots) byslot[(size_t)kv.second] = kv.first;
    printf("%s ", name);
    for (size_t i = 0; i < byslot.size(); i++) printf("%s%s", i ? "," : "", byslot[i].c_str());
    for (auto w : s.words) printf(" %lld", (long long)w);
    printf("\n");
}

int main()
{
    dump("q6", *q6_plan());
    dump("q1", *q1_plan());
    dump("q3", *q3_plan(10));
    dump("q18", *q18_plan(314, 100));
    dump("q9", *q9_plan("pink"));
    return 0;
}

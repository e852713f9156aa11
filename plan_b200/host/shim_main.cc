// shim_main.cc -- drives the C++ host shim the way cmd/tester drives the reference
// (`tester tpch1g --query_id N`, /root/reference/cmd/tester/main.go:69-72): builds the physical
// plan, pulls the root OperatorExec chunk by chunk (execOps, executor.go:151-188) and writes the
// result file text (executor_bench.go:215-241: '#' headline, tab separated rows) to stdout.
//   planhost_run <scale factor> <query id: 1|3|6|9|18>
#include <stdio.h>
#include <stdlib.h>

#include "tpch_plans.hpp"

using namespace planhost;

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <sf> <query_id>\n", argv[0]); return 2; }
    double sf = atof(argv[1]);
    int q = atoi(argv[2]);
    try {
        check(pg_init(0));
        pg_table *orders = nullptr, *lineitem = nullptr, *customer = nullptr;
        check(pg_tpch_orders_lineitem(sf, 0, pg_tpch_num_orders(sf), &orders, &lineitem));
        check(pg_tpch_customer(sf, 0, pg_tpch_num_customers(sf), &customer));
        std::map<std::string, pg_table *> tables = {{"lineitem", lineitem}, {"orders", orders}, {"customer", customer}};
        pg_table *part = nullptr, *supplier = nullptr, *partsupp = nullptr, *nation = nullptr;
        if (q == 9) {
            check(pg_tpch_part(sf, &part));
            check(pg_tpch_supplier(sf, &supplier));
            check(pg_tpch_partsupp(sf, &partsupp));
            check(pg_tpch_nation(&nation));
            tables["part"] = part; tables["supplier"] = supplier; tables["partsupp"] = partsupp; tables["nation"] = nation;
        }
        Op plan = q == 6 ? q6_plan() : q == 1 ? q1_plan() : q == 18 ? q18_plan() : q == 9 ? q9_plan() : q3_plan();
        GpuPipelineExec ex(plan, tables);
        ex.Init();
        fprintf(stderr, "%s\n", ex.Explain());
        bool head = false;
        size_t ncols = 0;
        {
            const PhysicalOperator *agg = plan.get();
            while (agg->Typ == POT_Limit || agg->Typ == POT_Order) agg = agg->Children[0].get();
            ncols = agg->Outputs.size();
        }
        fputc('#', stdout);
        for (size_t i = 1; i < ncols; i++) fputc('\t', stdout);
        fputc('\n', stdout);
        (void)head;
        for (;;) {
            Chunk out;
            OperatorResult r = ex.Execute(nullptr, &out);
            if (r == Done) break;
            out.SaveToFile(stdout);
        }
        ex.Close();
        pg_table_free(orders);
        pg_table_free(lineitem);
        pg_table_free(customer);
        pg_table_free(part);
        pg_table_free(supplier);
        pg_table_free(partsupp);
        pg_table_free(nation);
    } catch (const PlanError &e) {
        fprintf(stderr, "plangpu error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}

// gpu_exec.hpp -- C++ host shim above the C ABI: the operator the reference's buildOperatorExec
// would construct for an off-loadable subtree.
//   OperatorExec{Init,Execute,Close}  /root/reference/pkg/compute/executor_operator.go:52-56
//   OperatorResult                    /root/reference/pkg/compute/executor_operator.go:11-18
//   PhysicalOperator / Expr           /root/reference/pkg/compute/builder_physical_operator.go:49-66, expr.go:49-60
//   operator infos                    /root/reference/pkg/compute/operator_info.go:10-34
// Only pg_* entry points of include/plangpu.h are called; no CUDA or torch type appears here.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>

#include "chunk.hpp"

namespace planhost {

enum OperatorResult { InvalidOpResult = 0, NeedMoreInput = 1, haveMoreOutput = 2, Done = 3 };
enum POT { POT_Scan = 1, POT_Filter = 2, POT_Join = 3, POT_Agg = 4, POT_Project = 5, POT_Order = 6, POT_Limit = 7 };
enum ET { ET_Column = 0, ET_Func = 5, ET_Const = 7 };

struct Expr {
    ET Typ = ET_Column;
    LType DataTyp;
    std::vector<Expr> Children;
    int Side = 0, Idx = 0;            // ColRef
    int64_t I64 = 0;                  // ConstValue (integers, dates, unscaled decimals)
    double F64 = 0;                   // ConstValue (FLOAT / DOUBLE)
    std::string Str;                  // ConstValue (VARCHAR)
    std::string FunImpl;              // function name (function.go:89-128)
};

inline Expr col(int side, int idx, LType t) { Expr e; e.Typ = ET_Column; e.DataTyp = t; e.Side = side; e.Idx = idx; return e; }
inline Expr constI(int64_t v, LType t) { Expr e; e.Typ = ET_Const; e.DataTyp = t; e.I64 = v; return e; }
inline Expr constF(double v, LType t) { Expr e; e.Typ = ET_Const; e.DataTyp = t; e.F64 = v; return e; }
inline Expr constS(const std::string &s) { Expr e; e.Typ = ET_Const; e.DataTyp = LType::Varchar(); e.Str = s; return e; }
inline Expr func(const std::string &name, LType t, std::vector<Expr> ch) { Expr e; e.Typ = ET_Func; e.DataTyp = t; e.FunImpl = name; e.Children = std::move(ch); return e; }
inline Expr cast(Expr c, LType t) { return func("cast", t, {std::move(c)}); }

struct PhysicalOperator {
    POT Typ = POT_Scan;
    std::vector<Expr> Outputs, Filters;
    std::vector<std::shared_ptr<PhysicalOperator>> Children;
    // Info
    std::string Table;                           // ScanOpInfo
    int JoinTyp = PG_JOIN_INNER;                 // JoinOpInfo
    std::vector<Expr> OnConds;
    std::vector<Expr> Aggs, GroupBys;            // AggOpInfo (Aggs[i] = func("sum", result type, {arg}))
    std::vector<std::pair<Expr, bool>> OrderBys; // OrderOpInfo (column ref into the child outputs, descending)
    int64_t Limit = -1;                          // LimitOpInfo
};
typedef std::shared_ptr<PhysicalOperator> Op;

struct PlanError : std::runtime_error {
    int status;
    PlanError(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};
inline void check(int rc) { if (rc != PG_OK) throw PlanError(rc, pg_last_error()); }

// ------------------------------------------------------------- plan descriptor --
class Serializer {
public:
    std::vector<int64_t> words;
    std::map<std::string, int> slots;

    static int fn_id(const std::string &n)
    {
        static const std::map<std::string, int> m = {{"+", PG_FN_ADD}, {"-", PG_FN_SUB}, {"*", PG_FN_MUL}, {"/", PG_FN_DIV},
            {"=", PG_FN_EQ}, {"<>", PG_FN_NE}, {"<", PG_FN_LT}, {"<=", PG_FN_LE}, {">", PG_FN_GT}, {">=", PG_FN_GE}, {"in", PG_FN_IN},
            {"like", PG_FN_LIKE}, {"not like", PG_FN_NOT_LIKE}, {"extract", PG_FN_EXTRACT},
            {"and", PG_FN_AND}, {"or", PG_FN_OR}, {"not", PG_FN_NOT}, {"cast", PG_FN_CAST}};
        auto it = m.find(n);
        if (it == m.end()) throw PlanError(PG_EUNSUPPORTED, "function " + n + " cannot be off-loaded");
        return it->second;
    }
    static int agg_id(const std::string &n)
    {
        static const std::map<std::string, int> m = {{"sum", PG_AGG_SUM}, {"avg", PG_AGG_AVG}, {"count", PG_AGG_COUNT}, {"min", PG_AGG_MIN}, {"max", PG_AGG_MAX}};
        auto it = m.find(n);
        if (it == m.end()) throw PlanError(PG_EUNSUPPORTED, "aggregate " + n + " cannot be off-loaded");
        return it->second;
    }
    void ltype(const LType &t) { words.insert(words.end(), {t.Id, t.Width, t.Scale}); }
    int tokens(const Expr &e, std::vector<int64_t> &out)
    {
        if (e.Typ == ET_Column) { out.insert(out.end(), {PG_TK_COL, e.Side, e.Idx, e.DataTyp.Id, e.DataTyp.Width, e.DataTyp.Scale}); return 1; }
        if (e.Typ == ET_Const) {
            if (e.DataTyp.Id == PG_LT_VARCHAR) {
                out.insert(out.end(), {PG_TK_STR, (int64_t)e.Str.size()});
                for (size_t i = 0; i < e.Str.size(); i += 8) {
                    int64_t w = 0;
                    for (size_t b = 0; b < 8 && i + b < e.Str.size(); b++) w |= (int64_t)(uint8_t)e.Str[i + b] << (8 * b);
                    out.push_back(w);
                }
                return 1;
            }
            int64_t v = e.I64;
            if (e.DataTyp.Id == PG_LT_FLOAT || e.DataTyp.Id == PG_LT_DOUBLE) memcpy(&v, &e.F64, 8);
            out.insert(out.end(), {PG_TK_CONST, e.DataTyp.Id, e.DataTyp.Width, e.DataTyp.Scale, v});
            return 1;
        }
        int n = 0;
        for (auto &c : e.Children) n += tokens(c, out);
        out.insert(out.end(), {PG_TK_FUNC, fn_id(e.FunImpl), (int64_t)e.Children.size(), e.DataTyp.Id, e.DataTyp.Width, e.DataTyp.Scale});
        return n + 1;
    }
    void expr(const Expr &e)
    {
        std::vector<int64_t> t;
        int n = tokens(e, t);
        words.push_back(n);
        words.insert(words.end(), t.begin(), t.end());
    }
    void node(const PhysicalOperator &op)
    {
        switch (op.Typ) {
        case POT_Limit:
        case POT_Order: {
            const PhysicalOperator *order = op.Typ == POT_Limit ? op.Children[0].get() : &op;
            if (order->Typ != POT_Order || order->Children[0]->Typ != POT_Agg) throw PlanError(PG_EUNSUPPORTED, "only Limit<-Order<-Agg fuses");
            words.insert(words.end(), {PG_OP_TOPK, (int64_t)order->OrderBys.size()});
            for (auto &o : order->OrderBys) words.insert(words.end(), {o.first.Idx, o.second ? 1 : 0});
            words.push_back(op.Typ == POT_Limit ? op.Limit : -1);
            node(*order->Children[0]);
            break;
        }
        case POT_Scan:
            if (!slots.count(op.Table)) { int s = (int)slots.size(); slots[op.Table] = s; }
            words.insert(words.end(), {PG_OP_SCAN, slots[op.Table], (int64_t)op.Filters.size()});
            for (auto &f : op.Filters) expr(f);
            break;
        case POT_Filter:
            words.insert(words.end(), {PG_OP_FILTER, (int64_t)op.Filters.size()});
            for (auto &f : op.Filters) expr(f);
            node(*op.Children[0]);
            break;
        case POT_Join:
            words.insert(words.end(), {PG_OP_JOIN, op.JoinTyp, (int64_t)op.OnConds.size()});
            for (auto &c : op.OnConds) { expr(c.Children[0]); expr(c.Children[1]); }
            words.push_back((int64_t)op.Outputs.size());
            for (auto &o : op.Outputs) words.insert(words.end(), {o.Side, o.Idx});
            node(*op.Children[0]);
            node(*op.Children[1]);
            break;
        case POT_Agg:
            words.insert(words.end(), {PG_OP_AGG, (int64_t)op.GroupBys.size()});
            for (auto &g : op.GroupBys) expr(g);
            words.push_back((int64_t)op.Aggs.size());
            for (auto &a : op.Aggs) {
                words.push_back(agg_id(a.FunImpl));
                ltype(a.DataTyp);
                if (a.Children.empty()) words.push_back(0); else expr(a.Children[0]);
            }
            words.push_back((int64_t)op.Filters.size());
            for (auto &f : op.Filters) expr(f);
            words.push_back((int64_t)op.Outputs.size());
            for (auto &o : op.Outputs) words.insert(words.end(), {o.Side, o.Idx});
            node(*op.Children[0]);
            break;
        default: throw PlanError(PG_EUNSUPPORTED, "operator cannot be off-loaded");
        }
    }
    void plan(const PhysicalOperator &op)
    {
        words = {PG_DESC_MAGIC, PG_DESC_VERSION};
        node(op);
    }
};

// ------------------------------------------------------------------ executor --
struct OperatorExec {
    virtual ~OperatorExec() {}
    virtual void Init() = 0;
    virtual OperatorResult Execute(Chunk *input, Chunk *output) = 0;
    virtual void Close() = 0;
};

class GpuPipelineExec : public OperatorExec {
public:
    GpuPipelineExec(Op op, std::map<std::string, pg_table *> tables) : op_(std::move(op)), tables_(std::move(tables)) {}
    ~GpuPipelineExec() override { Close(); }

    void Init() override
    {
        Serializer s;
        s.plan(*op_);
        check(pg_plan_compile(s.words.data(), s.words.size(), &plan_));
        for (auto &kv : s.slots) {
            auto it = tables_.find(kv.first);
            if (it == tables_.end()) throw PlanError(PG_EINVAL, "no device table for " + kv.first);
            check(pg_plan_bind(plan_, kv.second, it->second));
        }
        check(pg_plan_prepare(plan_));     // PG_EUNSUPPORTED: the caller builds the stock executors
    }

    OperatorResult Execute(Chunk *, Chunk *output) override
    {
        if (!result_) {
            check(pg_plan_execute(plan_, &result_));
            check(pg_result_num_columns(result_, &ncols_));
        }
        int64_t n = 0;
        std::vector<const void *> cols((size_t)ncols_);
        check(pg_result_next(result_, DefaultVectorSize, &n, cols.data(), nullptr));
        output->Data.clear();
        output->Count = n;
        if (n == 0) return Done;
        const PhysicalOperator *agg = op_.get();
        while (agg->Typ == POT_Limit || agg->Typ == POT_Order) agg = agg->Children[0].get();
        for (int i = 0; i < ncols_; i++) {
            int32_t t, w, sc;
            check(pg_result_column_type(result_, i, &t, &w, &sc));
            Vector v;
            v.NativeType = t;
            v.Typ = (size_t)i < agg->Outputs.size() ? agg->Outputs[(size_t)i].DataTyp : LType();
            size_t esz = t == PG_T_HUGEINT || t == PG_T_DECIMAL128 || t == PG_T_VARCHAR ? 16 : (t == PG_T_INT32 || t == PG_T_DATE32) ? 4 : (t == PG_T_CHAR1 || t == PG_T_DICT8) ? 1 : 8;
            v.Data.assign((const uint8_t *)cols[(size_t)i], (const uint8_t *)cols[(size_t)i] + esz * (size_t)n);
            if (t == PG_T_DICT8) {
                int32_t nd = 0;
                const char *const *ents = nullptr;
                check(pg_result_column_dict(result_, i, &nd, &ents));
                for (int32_t k = 0; k < nd; k++) v.Dict.emplace_back(ents[k]);
            }
            if (t == PG_T_VARCHAR) {
                const pg_string *sv = (const pg_string *)cols[(size_t)i];
                for (int64_t r = 0; r < n; r++) v.Strings.emplace_back(sv[r].data, (size_t)sv[r].len);
            }
            output->Data.push_back(std::move(v));
        }
        return haveMoreOutput;
    }

    void Close() override
    {
        if (result_) { pg_result_free(result_); result_ = nullptr; }
        if (plan_) { pg_plan_free(plan_); plan_ = nullptr; }
    }
    const char *Explain() { return pg_plan_explain(plan_); }

private:
    Op op_;
    std::map<std::string, pg_table *> tables_;
    pg_plan *plan_ = nullptr;
    pg_result *result_ = nullptr;
    int ncols_ = 0;
};

}  // namespace planhost

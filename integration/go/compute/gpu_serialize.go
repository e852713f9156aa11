// gpu_serialize.go -- PhysicalOperator subtree -> the flat int64 plan descriptor of include/plangpu_desc.h.
//
// Drop into pkg/compute of daviszhen/plan (package compute).  Go twin of plan_b200/compute.py::serialize_plan,
// which feeds the same descriptor to the same library in every GPU test of this repository.
//
// Reference structures walked: PhysicalOperator{Typ, Outputs, Filters, Children, Info}
// (builder_physical_operator.go:49-66), AggOpInfo / JoinOpInfo / ScanOpInfo / OrderOpInfo / LimitOpInfo
// (operator_info.go:10-45), Expr{Typ, DataTyp, ColRef, ConstValue, Children, Info -> FunctionInfo.FunImpl}
// (expr.go:49-60,125-128), function names (function.go:89-128).
package compute

import (
	"fmt"
	"math"

	"github.com/daviszhen/plan/pkg/common"
)

// descriptor constants (include/plangpu_desc.h)
const (
	pgDescMagic   = 0x31504750
	pgDescVersion = 1

	pgOpScan, pgOpFilter, pgOpJoin, pgOpAgg, pgOpTopK, pgOpProject = 1, 2, 3, 4, 5, 6

	pgTkCol, pgTkConst, pgTkStr, pgTkFunc = 1, 2, 3, 4

	pgLtBoolean, pgLtInteger, pgLtBigint, pgLtDate, pgLtDecimal = 1, 2, 3, 4, 5
	pgLtFloat, pgLtDouble, pgLtVarchar, pgLtHugeint             = 6, 7, 8, 9
)

var pgFuncIDs = map[string]int64{
	FuncAdd: 1, FuncSubtract: 2, FuncMultiply: 3, FuncDivide: 4,
	FuncEqual: 10, FuncNotEqual: 11, FuncLess: 12, FuncLessEqual: 13, FuncGreater: 14, FuncGreaterEqual: 15,
	FuncIn: 16, FuncLike: 17, FuncNotLike: 18, FuncExtract: 19,
	FuncAnd: 20, FuncOr: 21, FuncNot: 22,
	FuncCase: 23, // Children = [ELSE, WHEN1, THEN1, ...] as bindCaseExpr builds them (builder_expr.go:95-113)
	FuncCast: 30,
}

var pgAggIDs = map[string]int64{"sum": 1, "avg": 2, "count": 3, "min": 4, "max": 5}

var pgJoinTypes = map[LOT_JoinType]int64{
	LOT_JoinTypeInner: 1, LOT_JoinTypeSEMI: 2, LOT_JoinTypeANTI: 3, LOT_JoinTypeMARK: 4, LOT_JoinTypeLeft: 5, LOT_JoinTypeAntiMARK: 6,
}

// errNotOffloadable makes newGpuPipelineExec return an error so that buildOperatorExec builds the stock executors.
type errNotOffloadable struct{ why string }

func (e errNotOffloadable) Error() string { return "gpu: not off-loadable: " + e.why }

func pgLType(t common.LType) ([]int64, error) {
	var id int64
	switch t.Id {
	case common.LTID_BOOLEAN:
		id = pgLtBoolean
	case common.LTID_INTEGER:
		id = pgLtInteger
	case common.LTID_BIGINT:
		id = pgLtBigint
	case common.LTID_DATE:
		id = pgLtDate
	case common.LTID_DECIMAL:
		id = pgLtDecimal
	case common.LTID_FLOAT:
		id = pgLtFloat
	case common.LTID_DOUBLE:
		id = pgLtDouble
	case common.LTID_VARCHAR:
		id = pgLtVarchar
	case common.LTID_HUGEINT:
		id = pgLtHugeint
	default:
		return nil, errNotOffloadable{fmt.Sprintf("logical type %v", t)}
	}
	return []int64{id, int64(t.Width), int64(t.Scale)}, nil
}

// planSerializer carries the table -> slot assignment and, per operator, the output lists that column references
// of its parent index.
type planSerializer struct {
	slots map[string]int // "db.table" -> bind slot
	scans []*PhysicalOperator
}

// colRef maps an Expr column reference to (side, idx) of the descriptor.  In the physical plan a reference to a child
// output has ColRef.table() < 0: child number = -table-1 (expr_exec.go:248-265); a reference >= 0 inside an aggregate
// addresses the aggregate's own state: table == op.Index -> group key, table == AggTag -> aggregate (expr.go:743-763).
func (ps *planSerializer) colRef(e *Expr, agg *PhysicalOperator) (side, idx int64, err error) {
	tab := int64(e.ColRef.table())
	col := int64(e.ColRef.column())
	if tab < 0 {
		return -tab - 1, col, nil
	}
	if agg != nil {
		switch uint64(tab) {
		case agg.Index:
			return 0, col, nil
		case agg.getAggTag():
			return 1, col, nil
		}
	}
	return 0, 0, errNotOffloadable{fmt.Sprintf("column reference (%d,%d) not resolvable", tab, col)}
}

func (ps *planSerializer) exprTokens(e *Expr, agg *PhysicalOperator, ntok *int64, out *[]int64) error {
	switch e.Typ {
	case ET_Column:
		side, idx, err := ps.colRef(e, agg)
		if err != nil {
			return err
		}
		lt, err := pgLType(e.DataTyp)
		if err != nil {
			return err
		}
		*out = append(*out, pgTkCol, side, idx)
		*out = append(*out, lt...)
		*ntok++
	case ET_Const:
		lt, err := pgLType(e.DataTyp)
		if err != nil {
			return err
		}
		switch e.ConstValue.Type {
		case ConstTypeString:
			b := []byte(e.ConstValue.String)
			*out = append(*out, pgTkStr, int64(len(b)))
			for i := 0; i < len(b); i += 8 {
				var w uint64
				for k := 0; k < 8 && i+k < len(b); k++ {
					w |= uint64(b[i+k]) << (8 * uint(k))
				}
				*out = append(*out, int64(w))
			}
		case ConstTypeFloat:
			*out = append(*out, pgTkConst)
			*out = append(*out, lt...)
			*out = append(*out, int64(math.Float64bits(e.ConstValue.Float)))
		case ConstTypeInteger:
			*out = append(*out, pgTkConst)
			*out = append(*out, lt...)
			*out = append(*out, e.ConstValue.Integer)
		case ConstTypeDate:
			d, err := parseDateDays(e.ConstValue.Date) // days since 1970-01-01, as the device stores DATE
			if err != nil {
				return errNotOffloadable{err.Error()}
			}
			*out = append(*out, pgTkConst)
			*out = append(*out, lt...)
			*out = append(*out, d)
		case ConstTypeDecimal:
			u, err := parseDecimalUnscaled(e.ConstValue.Decimal, e.DataTyp.Scale) // unscaled value at DataTyp.Scale
			if err != nil {
				return errNotOffloadable{err.Error()}
			}
			*out = append(*out, pgTkConst)
			*out = append(*out, lt...)
			*out = append(*out, u)
		case ConstTypeBoolean:
			v := int64(0)
			if e.ConstValue.Boolean {
				v = 1
			}
			*out = append(*out, pgTkConst)
			*out = append(*out, lt...)
			*out = append(*out, v)
		case ConstTypeNull:
			*out = append(*out, pgTkConst, 0, 0, 0, 0) // ltype 0 = NULL (a CASE without ELSE)
		default:
			return errNotOffloadable{fmt.Sprintf("constant type %v", e.ConstValue.Type)}
		}
		*ntok++
	case ET_Func:
		fi := e.GetFuncInfo()
		id, ok := pgFuncIDs[fi.FunImpl.Name()]
		if !ok {
			return errNotOffloadable{"function " + fi.FunImpl.Name()}
		}
		for _, c := range e.Children {
			if err := ps.exprTokens(c, agg, ntok, out); err != nil {
				return err
			}
		}
		lt, err := pgLType(e.DataTyp)
		if err != nil {
			return err
		}
		*out = append(*out, pgTkFunc, id, int64(len(e.Children)))
		*out = append(*out, lt...)
		*ntok++
	default:
		return errNotOffloadable{fmt.Sprintf("expression type %v", e.Typ)}
	}
	return nil
}

// expr := ntokens token*   (postfix)
func (ps *planSerializer) exprWords(e *Expr, agg *PhysicalOperator, w *[]int64) error {
	var toks []int64
	var n int64
	if err := ps.exprTokens(e, agg, &n, &toks); err != nil {
		return err
	}
	*w = append(*w, n)
	*w = append(*w, toks...)
	return nil
}

func (ps *planSerializer) node(op *PhysicalOperator, w *[]int64) error {
	switch op.Typ {
	case POT_Limit, POT_Order:
		// Limit <- Order <- Agg (or Order <- Agg) with ORDER BY on plain output columns fuses into PG_OP_TOPK
		order, limit := op, int64(-1)
		if op.Typ == POT_Limit {
			if len(op.Children) != 1 || op.Children[0].Typ != POT_Order {
				return errNotOffloadable{"LIMIT without ORDER BY"}
			}
			order = op.Children[0]
			li := op.Info.(*LimitOpInfo)
			if li.Offset != nil || li.Limit == nil || li.Limit.Typ != ET_Const {
				return errNotOffloadable{"LIMIT with OFFSET or a computed count"}
			}
			limit = li.Limit.ConstValue.Integer
		}
		if len(order.Children) != 1 || order.Children[0].Typ != POT_Agg {
			return errNotOffloadable{"ORDER BY above something else than an aggregate"}
		}
		obs := order.Info.(*OrderOpInfo).OrderBys
		*w = append(*w, pgOpTopK, int64(len(obs)))
		for _, ob := range obs {
			key := ob.Children[0]
			if ob.Typ != ET_Orderby || key.Typ != ET_Column || int64(key.ColRef.table()) >= 0 {
				return errNotOffloadable{"ORDER BY key is not an output column of the aggregate"}
			}
			desc := int64(0)
			if ob.GetOrderByInfo().Desc {
				desc = 1
			}
			*w = append(*w, int64(key.ColRef.column()), desc)
		}
		*w = append(*w, limit)
		return ps.node(order.Children[0], w)
	case POT_Scan:
		si := op.Info.(*ScanOpInfo)
		if si.ScanTyp != ScanTypeTable {
			return errNotOffloadable{"scan of something else than a stored table"}
		}
		key := si.Database + "." + si.Table
		slot, ok := ps.slots[key]
		if !ok {
			slot = len(ps.slots)
			ps.slots[key] = slot
			ps.scans = append(ps.scans, op)
		}
		*w = append(*w, pgOpScan, int64(slot), int64(len(op.Filters)))
		for _, f := range op.Filters {
			if err := ps.exprWords(f, nil, w); err != nil {
				return err
			}
		}
		return nil
	case POT_Project:
		// root of a ROW-EMITTING pipeline (rows.cu): Project <- [Filter]* <- (Scan | Join(scan, scan)).  A Project right
		// under an aggregate never gets here: POT_Agg inlines it below.
		*w = append(*w, pgOpProject, int64(len(op.Projects)))
		for _, e := range op.Projects {
			if err := ps.exprWords(e, nil, w); err != nil {
				return err
			}
		}
		return ps.node(op.Children[0], w)
	case POT_Filter:
		*w = append(*w, pgOpFilter, int64(len(op.Filters)))
		for _, f := range op.Filters {
			if err := ps.exprWords(f, nil, w); err != nil {
				return err
			}
		}
		return ps.node(op.Children[0], w)
	case POT_Join:
		ji := op.Info.(*JoinOpInfo)
		jt, ok := pgJoinTypes[ji.JoinTyp]
		if !ok {
			return errNotOffloadable{fmt.Sprintf("join type %v", ji.JoinTyp)}
		}
		*w = append(*w, pgOpJoin, jt, int64(len(ji.OnConds)))
		for _, c := range ji.OnConds { // every condition is `=`(left key expr, right key expr) (join_table.go:85-120)
			if c.Typ != ET_Func || c.GetFuncInfo().FunImpl.Name() != FuncEqual || len(c.Children) != 2 {
				return errNotOffloadable{"join condition other than an equality"}
			}
			if err := ps.exprWords(c.Children[0], nil, w); err != nil {
				return err
			}
			if err := ps.exprWords(c.Children[1], nil, w); err != nil {
				return err
			}
		}
		*w = append(*w, int64(len(op.Outputs)))
		for _, o := range op.Outputs { // plain references to child outputs (executor_join.go:237-264)
			if o.Typ != ET_Column || int64(o.ColRef.table()) >= 0 {
				return errNotOffloadable{"join output is not a child column"}
			}
			*w = append(*w, -int64(o.ColRef.table())-1, int64(o.ColRef.column()))
		}
		if err := ps.node(op.Children[0], w); err != nil {
			return err
		}
		return ps.node(op.Children[1], w)
	case POT_Agg:
		ai := op.Info.(*AggOpInfo)
		child := op.Children[0]
		// a Project between the aggregate and its input is inlined by substituting its expressions
		groups, aggs := ai.GroupBys, ai.Aggs
		if child.Typ == POT_Project {
			groups = inlineProject(groups, child)
			aggs = inlineProject(aggs, child)
			child = child.Children[0]
		}
		*w = append(*w, pgOpAgg, int64(len(groups)))
		for _, g := range groups {
			if err := ps.exprWords(g, nil, w); err != nil {
				return err
			}
		}
		*w = append(*w, int64(len(aggs)))
		for _, a := range aggs {
			fi := a.GetFuncInfo()
			id, ok := pgAggIDs[fi.FunImpl.Name()]
			if !ok || fi.FunImpl.IsDistinct() {
				return errNotOffloadable{"aggregate " + fi.FunImpl.Name()}
			}
			lt, err := pgLType(a.DataTyp) // the aggregate's RESULT type (function_aggr.go:48-103)
			if err != nil {
				return err
			}
			*w = append(*w, id)
			*w = append(*w, lt...)
			if len(a.Children) == 0 {
				*w = append(*w, 0) // count(*)
			} else if err := ps.exprWords(a.Children[0], nil, w); err != nil {
				return err
			}
		}
		*w = append(*w, int64(len(op.Filters))) // HAVING
		for _, f := range op.Filters {
			if err := ps.exprWords(f, op, w); err != nil {
				return err
			}
		}
		*w = append(*w, int64(len(op.Outputs)))
		for _, o := range op.Outputs {
			if o.Typ != ET_Column {
				return errNotOffloadable{"aggregate output is an expression"}
			}
			kind, idx, err := ps.colRef(o, op)
			if err != nil {
				return err
			}
			if int64(o.ColRef.table()) < 0 {
				return errNotOffloadable{"aggregate output refers to a child column (referChildren, executor_aggr.go:92-99)"}
			}
			*w = append(*w, kind, idx)
		}
		return ps.node(child, w)
	}
	return errNotOffloadable{fmt.Sprintf("operator %v", op.Typ)}
}

// inlineProject rewrites references to the outputs of a POT_Project by the project's own expressions.
func inlineProject(exprs []*Expr, proj *PhysicalOperator) []*Expr {
	var sub func(e *Expr) *Expr
	sub = func(e *Expr) *Expr {
		if e.Typ == ET_Column && int64(e.ColRef.table()) == -1 {
			return proj.Projects[e.ColRef.column()]
		}
		c := *e
		c.Children = make([]*Expr, len(e.Children))
		for i, ch := range e.Children {
			c.Children[i] = sub(ch)
		}
		return &c
	}
	out := make([]*Expr, len(exprs))
	for i, e := range exprs {
		out[i] = sub(e)
	}
	return out
}

// serializePlan returns the descriptor, the scans to bind (slot i = scans[i]) and an error that means "build the stock
// executors" when the subtree has a shape the library does not take.
func serializePlan(op *PhysicalOperator) (desc []int64, scans []*PhysicalOperator, err error) {
	ps := &planSerializer{slots: map[string]int{}}
	desc = []int64{pgDescMagic, pgDescVersion}
	if err = ps.node(op, &desc); err != nil {
		return nil, nil, err
	}
	return desc, ps.scans, nil
}

// parseDateDays: 'YYYY-MM-DD' -> days since 1970-01-01 (the reference parses the same literal in
// tryCastVarcharToDate, function_cast.go:425-447).
func parseDateDays(s string) (int64, error) {
	var y, m, d int
	if _, err := fmt.Sscanf(s, "%d-%d-%d", &y, &m, &d); err != nil {
		return 0, fmt.Errorf("date literal %q", s)
	}
	dt := common.Date{Year: int32(y), Month: int32(m), Day: int32(d)}
	return dt.ToDate().Unix() / 86400, nil
}

// parseDecimalUnscaled: decimal literal -> unscaled integer at `scale` (exact; more fractional digits than `scale`
// are refused rather than rounded).
func parseDecimalUnscaled(s string, scale int) (int64, error) {
	neg := false
	if len(s) > 0 && (s[0] == '-' || s[0] == '+') {
		neg = s[0] == '-'
		s = s[1:]
	}
	var v int64
	frac := -1
	for _, ch := range s {
		switch {
		case ch == '.':
			if frac >= 0 {
				return 0, fmt.Errorf("decimal literal %q", s)
			}
			frac = 0
		case ch >= '0' && ch <= '9':
			if v > (math.MaxInt64-9)/10 {
				return 0, fmt.Errorf("decimal literal %q overflows", s)
			}
			v = v*10 + int64(ch-'0')
			if frac >= 0 {
				frac++
			}
		default:
			return 0, fmt.Errorf("decimal literal %q", s)
		}
	}
	if frac < 0 {
		frac = 0
	}
	for ; frac < scale; frac++ {
		if v > math.MaxInt64/10 {
			return 0, fmt.Errorf("decimal literal %q overflows", s)
		}
		v *= 10
	}
	if frac > scale {
		return 0, fmt.Errorf("decimal literal %q has more than %d fractional digits", s, scale)
	}
	if neg {
		v = -v
	}
	return v, nil
}

// executor_gpu.go -- the OperatorExec that runs a fusable subtree on the GPU through libplangpu.
//
// Drop into pkg/compute of daviszhen/plan (package compute).  The one change to existing code is the arm at the
// top of buildOperatorExec (executor.go:305-350), BEFORE the children are built:
//
//	if cfg.Gpu.Enable {        // len(cfg.Gpu.Devices) > 1: newMultiGpuPipelineExec (executor_gpu_multi.go) instead
//		if ex, err := newGpuPipelineExec(op, cfg, txn); err == nil {
//			if err = ex.Init(); err == nil {
//				return ex, nil          // the whole subtree runs on the device
//			}
//			ex.Close()                  // PG_EUNSUPPORTED etc.: fall through to the stock executors
//		}
//	}
//
// (util.Config gains `Gpu struct{ Enable bool; Device int; Devices []int; ReplicateBelow int }`, pkg/util/config.go:56-59.)
//
// Interface implemented: OperatorExec{Init, Execute, Close} (executor_operator.go:52-56) with the results of
// executor_operator.go:11-18.  Parents (Project / Order / Limit executors, Runner.Execute executor.go:242-296)
// pull <= util.DefaultVectorSize rows per Execute exactly as from the stock aggExecutor.
//
// Go twin of plan_b200/compute.py::gpuPipelineExec and plan_b200/host/gpu_exec.hpp, which run the same sequence in
// this repository's tests.  Not compiled here (no Go toolchain in the image).
package compute

import (
	"fmt"
	"sync"
	"time"
	"unsafe"

	"github.com/daviszhen/plan/pkg/chunk"
	"github.com/daviszhen/plan/pkg/common"
	"github.com/daviszhen/plan/pkg/storage"
	"github.com/daviszhen/plan/pkg/util"

	"github.com/daviszhen/plan/pkg/compute/plangpu"
)

// ---------------------------------------------------------------- device column cache --

// deviceTables caches sealed device tables per (database.table, storage version): the second query over a table
// binds the resident copy instead of draining the scan again (SURVEY 8f-2; executor_scan.go:158-223 re-reads
// storage for every query).
var deviceTables = struct {
	sync.Mutex
	m map[string]*deviceTable
}{m: map[string]*deviceTable{}}

type deviceTable struct {
	tab     *plangpu.Table
	version uint64
	cols    []plangpu.ColDesc
}

// columnEncoding picks the device encoding of a scan output column.
func columnEncoding(name string, t common.LType) (plangpu.ColDesc, error) {
	d := plangpu.ColDesc{Name: name, Width: int32(t.Width), Scale: int32(t.Scale)}
	switch t.Id {
	case common.LTID_INTEGER:
		d.Type = plangpu.TInt32
	case common.LTID_BIGINT:
		d.Type = plangpu.TInt64
	case common.LTID_DATE:
		d.Type = plangpu.TDate32
	case common.LTID_DECIMAL:
		if t.Width > 19 {
			return d, errNotOffloadable{"DECIMAL wider than 19 digits: " + name}
		}
		d.Type = plangpu.TDecimal64
	case common.LTID_DOUBLE:
		d.Type = plangpu.TFloat64
	case common.LTID_VARCHAR:
		if t.Width == 1 {
			d.Type = plangpu.TChar1
		} else {
			d.Type = plangpu.TDict8 // demoted to TVarchar by the ingest when more than 256 distinct values turn up
		}
	default:
		return d, errNotOffloadable{fmt.Sprintf("column %s of type %v", name, t)}
	}
	return d, nil
}

// chunkStage is the flattened copy of ONE scan chunk: per column a narrow buffer + frame of reference.
type chunkStage struct {
	n    int
	cols []plangpu.ColBuf
	keep [][]byte // owners of the buffers
}

// flattenVector normalises a vector (FLAT / CONST / DICT via ToUnifiedFormat, vector_format.go:64-97) into the
// device-native encoding at the narrowest width the chunk's value range allows (pg_colbuf: value = base + stored).
// DECIMALs are read through the accessors the reference itself uses (Coef / Scale / IsNeg, chunk/hash.go:144-152).
func flattenVector(vec *chunk.Vector, count int, d *plangpu.ColDesc, dict map[string]uint8, dictList *[]string) (plangpu.ColBuf, []byte, error) {
	var uni chunk.UnifiedFormat
	vec.ToUnifiedFormat(count, &uni)
	vals := make([]int64, count)
	var valid []byte
	markNull := func(i int) {
		if valid == nil {
			valid = make([]byte, (count+7)/8)
			for k := range valid {
				valid[k] = 0xff
			}
		}
		valid[i>>3] &^= 1 << (uint(i) & 7)
	}
	switch d.Type {
	case plangpu.TInt32:
		src := chunk.GetSliceInPhyFormatUnifiedFormat[int32](&uni)
		for i := 0; i < count; i++ {
			j := uni.Sel.GetIndex(i)
			if !uni.Mask.RowIsValid(uint64(j)) {
				markNull(i)
				continue
			}
			vals[i] = int64(src[j])
		}
	case plangpu.TInt64:
		src := chunk.GetSliceInPhyFormatUnifiedFormat[int64](&uni)
		for i := 0; i < count; i++ {
			j := uni.Sel.GetIndex(i)
			if !uni.Mask.RowIsValid(uint64(j)) {
				markNull(i)
				continue
			}
			vals[i] = src[j]
		}
	case plangpu.TDate32:
		src := chunk.GetSliceInPhyFormatUnifiedFormat[common.Date](&uni)
		for i := 0; i < count; i++ {
			j := uni.Sel.GetIndex(i)
			if !uni.Mask.RowIsValid(uint64(j)) {
				markNull(i)
				continue
			}
			vals[i] = src[j].ToDate().Unix() / 86400 // days since 1970-01-01
		}
	case plangpu.TDecimal64:
		src := chunk.GetSliceInPhyFormatUnifiedFormat[common.Decimal](&uni)
		for i := 0; i < count; i++ {
			j := uni.Sel.GetIndex(i)
			if !uni.Mask.RowIsValid(uint64(j)) {
				markNull(i)
				continue
			}
			dec := src[j].Decimal
			u := int64(dec.Coef())
			for s := dec.Scale(); s < int(d.Scale); s++ { // value-preserving rescale to the column's declared scale
				u *= 10
			}
			if dec.Scale() > int(d.Scale) {
				return plangpu.ColBuf{}, nil, errNotOffloadable{"DECIMAL value with a larger scale than its column"}
			}
			if dec.IsNeg() {
				u = -u
			}
			vals[i] = u
		}
	case plangpu.TChar1, plangpu.TDict8:
		src := chunk.GetSliceInPhyFormatUnifiedFormat[common.String](&uni)
		out := make([]byte, count)
		for i := 0; i < count; i++ {
			j := uni.Sel.GetIndex(i)
			if !uni.Mask.RowIsValid(uint64(j)) {
				markNull(i)
				continue
			}
			if d.Type == plangpu.TChar1 {
				if src[j].Len > 0 {
					out[i] = src[j].DataSlice()[0]
				}
				continue
			}
			s := src[j].String()
			code, ok := dict[s]
			if !ok {
				if len(*dictList) == 256 {
					return plangpu.ColBuf{}, nil, errNotOffloadable{"more than 256 distinct strings in " + d.Name}
				}
				code = uint8(len(*dictList))
				dict[s] = code
				*dictList = append(*dictList, s)
			}
			out[i] = code
		}
		return plangpu.ColBuf{Data: unsafe.Pointer(&out[0]), Width: 0, Valid: valid}, out, nil
	case plangpu.TFloat64:
		src := chunk.GetSliceInPhyFormatUnifiedFormat[float64](&uni)
		out := make([]byte, 8*count)
		dst := unsafe.Slice((*float64)(unsafe.Pointer(&out[0])), count)
		for i := 0; i < count; i++ {
			j := uni.Sel.GetIndex(i)
			if !uni.Mask.RowIsValid(uint64(j)) {
				markNull(i)
				continue
			}
			dst[i] = src[j]
		}
		return plangpu.ColBuf{Data: unsafe.Pointer(&out[0]), Width: 0, Valid: valid}, out, nil
	default:
		return plangpu.ColBuf{}, nil, errNotOffloadable{"column encoding"}
	}
	// integer family: frame of reference at the narrowest width of [min, max] of this chunk
	lo, hi := vals[0], vals[0]
	for _, v := range vals[1:] {
		if v < lo {
			lo = v
		}
		if v > hi {
			hi = v
		}
	}
	span := uint64(hi - lo)
	var out []byte
	var width int32
	switch {
	case span <= 0xff:
		width = 1
		out = make([]byte, count)
		for i, v := range vals {
			out[i] = byte(v - lo)
		}
	case span <= 0xffff:
		width = 2
		out = make([]byte, 2*count)
		dst := unsafe.Slice((*uint16)(unsafe.Pointer(&out[0])), count)
		for i, v := range vals {
			dst[i] = uint16(v - lo)
		}
	case span <= 0x7fffffff:
		width = 4
		out = make([]byte, 4*count)
		dst := unsafe.Slice((*int32)(unsafe.Pointer(&out[0])), count)
		for i, v := range vals {
			dst[i] = int32(v - lo)
		}
	default:
		width, lo = 8, 0
		out = make([]byte, 8*count)
		copy(unsafe.Slice((*int64)(unsafe.Pointer(&out[0])), count), vals)
	}
	return plangpu.ColBuf{Data: unsafe.Pointer(&out[0]), Width: width, Base: lo, Valid: valid}, out, nil
}

// ingest drains a scan once (scanExecutor without its pushed-down filters: the filters travel in the plan and run
// on the device), flattening every 2048-row chunk and appending it (pg_table_append_cols gathers the small appends
// in pinned staging and copies asynchronously).  Dictionary columns need their dictionary at pg_table_create, so the
// flattened chunks are kept until the scan is drained and appended then.
func ingest(scanOp *PhysicalOperator, cfg *util.Config, txn *storage.Txn) (*deviceTable, error) {
	si := scanOp.Info.(*ScanOpInfo)
	descs, stages, err := drainScan(scanOp, cfg, txn)
	if err != nil {
		return nil, err
	}
	tab, err := plangpu.NewTable(si.Database+"."+si.Table, descs)
	if err != nil {
		return nil, err
	}
	for _, st := range stages {
		if err = tab.Append(st.n, st.cols); err != nil {
			tab.Free()
			return nil, err
		}
	}
	if err = tab.Seal(0); err != nil {
		tab.Free()
		return nil, err
	}
	return &deviceTable{tab: tab, cols: descs}, nil
}

// drainScan runs the scan to completion and returns the column descriptors (dictionaries filled in) with every
// 2048-row chunk flattened to narrow column buffers, in storage order.
func drainScan(scanOp *PhysicalOperator, cfg *util.Config, txn *storage.Txn) ([]plangpu.ColDesc, []chunkStage, error) {
	si := scanOp.Info.(*ScanOpInfo)
	plain := *scanOp
	plain.Filters = nil
	scan, err := newScanExecutor(&plain, cfg, txn, nil)
	if err != nil {
		return nil, nil, err
	}
	if err = scan.Init(); err != nil {
		return nil, nil, err
	}
	defer scan.Close()
	descs := make([]plangpu.ColDesc, len(scanOp.Outputs))
	dicts := make([]map[string]uint8, len(descs))
	dictLists := make([][]string, len(descs))
	for i, out := range scanOp.Outputs {
		if descs[i], err = columnEncoding(si.Columns[i], out.DataTyp); err != nil {
			return nil, nil, err
		}
		dicts[i] = map[string]uint8{}
	}
	var stages []chunkStage
	for {
		ch := &chunk.Chunk{}
		res, err := scan.Execute(nil, ch)
		if err != nil {
			return nil, nil, err
		}
		if res == Done {
			break
		}
		n := ch.Card()
		if n == 0 {
			continue
		}
		st := chunkStage{n: n}
		for i := range descs {
			cb, owner, err := flattenVector(ch.Data[i], n, &descs[i], dicts[i], &dictLists[i])
			if err != nil {
				return nil, nil, err
			}
			st.cols = append(st.cols, cb)
			st.keep = append(st.keep, owner)
		}
		stages = append(stages, st)
	}
	for i := range descs {
		descs[i].Dict = dictLists[i]
	}
	return descs, stages, nil
}

// deviceTableFor returns the cached device copy of the scan's table or ingests it.
func deviceTableFor(scanOp *PhysicalOperator, cfg *util.Config, txn *storage.Txn) (*deviceTable, error) {
	si := scanOp.Info.(*ScanOpInfo)
	key := si.Database + "." + si.Table
	ver := si.TableEnt.GetStorage().Version() // bumped by every committed write: a stale copy is never bound
	deviceTables.Lock()
	defer deviceTables.Unlock()
	if dt, ok := deviceTables.m[key]; ok {
		if dt.version == ver {
			return dt, nil
		}
		dt.tab.Free()
		delete(deviceTables.m, key)
	}
	dt, err := ingest(scanOp, cfg, txn)
	if err != nil {
		return nil, err
	}
	dt.version = ver
	deviceTables.m[key] = dt
	return dt, nil
}

// ---------------------------------------------------------------- the executor --

type gpuPipelineExec struct {
	op   *PhysicalOperator
	cfg  *util.Config
	txn  *storage.Txn
	plan *plangpu.Plan
	res  *plangpu.Result
}

func newGpuPipelineExec(op *PhysicalOperator, cfg *util.Config, txn *storage.Txn) (*gpuPipelineExec, error) {
	switch op.Typ {
	case POT_Agg, POT_Order, POT_Limit:
		return &gpuPipelineExec{op: op, cfg: cfg, txn: txn}, nil
	case POT_Project, POT_Filter, POT_Join, POT_Scan:
		// row-emitting pipelines (rows.cu): the library returns the operator's output ROWS in <= 2048-row chunks.
		// serializePlan / Prepare refuse what it does not take (deeper join trees, VARCHAR expressions ...).
		return &gpuPipelineExec{op: op, cfg: cfg, txn: txn}, nil
	}
	return nil, errNotOffloadable{"subtree root is neither an aggregate (optionally under Order / Limit) nor a row-emitting Project / Filter / Join / Scan"}
}

// Init serialises the subtree, binds device tables and lets the library choose its kernels.  An error (including
// PG_EUNSUPPORTED from Prepare) makes buildOperatorExec build the stock executors: plan selection, not a fallback
// at run time.
func (e *gpuPipelineExec) Init() error {
	desc, scans, err := serializePlan(e.op)
	if err != nil {
		return err
	}
	if e.plan, err = plangpu.Compile(desc); err != nil {
		return err
	}
	for slot, scanOp := range scans {
		dt, err := deviceTableFor(scanOp, e.cfg, e.txn)
		if err != nil {
			return err
		}
		if err = e.plan.Bind(slot, dt.tab); err != nil {
			return err
		}
	}
	return e.plan.Prepare()
}

func (e *gpuPipelineExec) Execute(input, output *chunk.Chunk) (OperatorResult, error) {
	ensureOutputChunk(e.op, output) // executor.go:201-210
	var err error
	if e.res == nil {
		if e.res, err = e.plan.Execute(); err != nil {
			// PG_EOVERFLOW is the class of faults the reference raises as a panic inside the executor
			// (function_operator_binary.go:134-140) and turns into an error in execQuery (executor_bench.go:184-189)
			return InvalidOpResult, err
		}
	}
	n, cols, valid, err := e.res.Next(util.DefaultVectorSize)
	if err != nil {
		return InvalidOpResult, err
	}
	if n == 0 {
		return Done, nil
	}
	for i, out := range e.op.Outputs {
		typ, _, scale := e.res.ColumnType(i)
		if err = fillVector(output.Data[i], out.DataTyp, typ, scale, cols[i], valid[i], n, e.res, i); err != nil {
			return InvalidOpResult, err
		}
	}
	output.SetCard(n)
	return haveMoreOutput, nil
}

func (e *gpuPipelineExec) Close() error {
	if e.res != nil {
		e.res.Free()
		e.res = nil
	}
	if e.plan != nil {
		e.plan.Free()
		e.plan = nil
	}
	return nil
}

// ---------------------------------------------------------------- results -> vectors --

type pgDecimal struct {
	Coef  uint64
	Scale int32
	Neg   uint32
}
type pgHugeint struct {
	Lower uint64
	Upper int64
}
type pgString struct { // field order of common.String (string.go:10-13)
	Len  int64
	Data unsafe.Pointer
}

func bitSet(bits *byte, i int) bool {
	return bits == nil || (*(*byte)(unsafe.Add(unsafe.Pointer(bits), i>>3))>>(uint(i)&7))&1 != 0
}

// fillVector writes one native result column into a FLAT vector of the operator's output type, through the same
// entry point the reference uses to load values (Vector.SetValue, chunk/vector.go:188-279): INT32 / INT64 / DOUBLE
// directly, DATE as {Year, Month, Day}, HUGEINT as {I64: upper, I64_1: lower}, DECIMAL through
// decimal.NewFromInt64(whole, frac, scale) -- the constructor of vector.go:256-263.
func fillVector(vec *chunk.Vector, lt common.LType, typ, scale int32, data unsafe.Pointer, valid *byte, n int, res *plangpu.Result, col int) error {
	var dict []string
	if typ == plangpu.TDict8 {
		dict = res.Dict(col)
	}
	for i := 0; i < n; i++ {
		val := &chunk.Value{Typ: lt}
		if !bitSet(valid, i) {
			val.IsNull = true
			vec.SetValue(i, val)
			continue
		}
		switch typ {
		case plangpu.TInt32:
			val.I64 = int64(*(*int32)(unsafe.Add(data, 4*i)))
		case plangpu.TInt64:
			val.I64 = *(*int64)(unsafe.Add(data, 8*i))
		case plangpu.TFloat64:
			val.F64 = *(*float64)(unsafe.Add(data, 8*i))
		case plangpu.TDate32:
			d := time.Unix(int64(*(*int32)(unsafe.Add(data, 4*i)))*86400, 0).UTC()
			val.I64, val.I64_1, val.I64_2 = int64(d.Year()), int64(d.Month()), int64(d.Day())
		case plangpu.TBool:
			val.Bool = *(*byte)(unsafe.Add(data, i)) != 0 // MARK columns, projected predicates (vector.go SetValue LTID_BOOLEAN)
		case plangpu.TChar1:
			val.Str = string([]byte{*(*byte)(unsafe.Add(data, i))})
		case plangpu.TDict8:
			code := int(*(*byte)(unsafe.Add(data, i)))
			if code >= len(dict) {
				return fmt.Errorf("gpu: dictionary code %d out of range", code)
			}
			val.Str = dict[code]
		case plangpu.TVarchar:
			s := (*pgString)(unsafe.Add(data, 16*i))
			val.Str = string(unsafe.Slice((*byte)(s.Data), int(s.Len))) // SetValue copies into C memory the vector owns
		case plangpu.THugeint:
			h := (*pgHugeint)(unsafe.Add(data, 16*i))
			val.I64, val.I64_1 = h.Upper, int64(h.Lower) // vector.go:272-275: Upper = I64, Lower = uint64(I64_1)
		case plangpu.TDecimal128:
			d := (*pgDecimal)(unsafe.Add(data, 16*i))
			// value = (-1)^neg * coef * 10^-scale, coef <= 10^19-1: split at the scale into (whole, frac)
			p := uint64(1)
			for k := int32(0); k < d.Scale; k++ {
				p *= 10
			}
			whole, frac := int64(d.Coef/p), int64(d.Coef%p)
			if d.Neg != 0 {
				whole, frac = -whole, -frac
			}
			// The vector's declared scale (lt.Scale) is what SetValue passes to NewFromInt64; the result scale can be
			// smaller (trailing zeros trimmed by the library exactly as govalues does), never larger.
			for k := d.Scale; k < int32(lt.Scale); k++ {
				frac *= 10
			}
			val.I64, val.I64_1 = whole, frac
		default:
			return fmt.Errorf("gpu: result column type %d", typ)
		}
		vec.SetValue(i, val)
	}
	_ = scale
	return nil
}

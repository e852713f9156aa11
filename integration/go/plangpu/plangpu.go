// Package plangpu binds the C ABI of libplangpu (include/plangpu.h) with cgo.
//
// Where it plugs in: pkg/compute/executor.go:305-350 (buildOperatorExec) builds a gpuPipelineExec
// (../compute/executor_gpu.go) for a fusable subtree; that executor talks to the GPU only through this
// package.  The reference already needs cgo (pkg/util/mem.go, pkg/storage/mem_buffer.go, AGENTS.md:25).
//
// cgo pointer rules honoured here:
//   - no Go pointer to memory that itself holds Go pointers crosses the boundary: the per-column pointer
//     tables handed to pg_table_append_cols / pg_result_next are allocated with C.malloc;
//   - the column buffers themselves are Go memory (util.GAlloc = make([]byte), pkg/util/mem.go:29-39); they are
//     pinned with runtime.Pinner for the duration of the call, and the library copies them before it returns;
//   - the library never calls back into Go and never retains a pointer after a call returns.
//
// NOT compiled in this repository's image (no Go toolchain: `go version` = command not found, here and on the
// GPU box); the same call sequence runs in every GPU test through plan_b200/_lib.py + compute.py (ctypes) and
// plan_b200/host/gpu_exec.hpp (C++).
package plangpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../plan_b200 -lplangpu -lcudart -ldl
#include <stdlib.h>
#include <string.h>
#include "plangpu.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Status codes (pg_status).
const (
	OK           = int(C.PG_OK)
	EInval       = int(C.PG_EINVAL)
	ENoMem       = int(C.PG_ENOMEM)
	ECuda        = int(C.PG_ECUDA)
	ENccl        = int(C.PG_ENCCL)
	EOverflow    = int(C.PG_EOVERFLOW)
	EUnsupported = int(C.PG_EUNSUPPORTED)
	EState       = int(C.PG_ESTATE)
)

// Column encodings (pg_type).
const (
	TInt32      = int32(C.PG_T_INT32)
	TInt64      = int32(C.PG_T_INT64)
	TDate32     = int32(C.PG_T_DATE32)
	TDecimal64  = int32(C.PG_T_DECIMAL64)
	TChar1      = int32(C.PG_T_CHAR1)
	TDict8      = int32(C.PG_T_DICT8)
	TFloat64    = int32(C.PG_T_FLOAT64)
	THugeint    = int32(C.PG_T_HUGEINT)
	TDecimal128 = int32(C.PG_T_DECIMAL128)
	TVarchar    = int32(C.PG_T_VARCHAR)
	TBool       = int32(C.PG_T_BOOL)
)

// Error carries the pg_status and the library's message for the calling thread.
type Error struct {
	Code int
	Msg  string
}

func (e *Error) Error() string { return fmt.Sprintf("plangpu status %d: %s", e.Code, e.Msg) }

// Unsupported reports PG_EUNSUPPORTED: the caller builds the stock executors for the subtree.
func Unsupported(err error) bool {
	e, ok := err.(*Error)
	return ok && e.Code == EUnsupported
}

// check must run on the OS thread that made the call (pg_last_error is thread local): every exported
// function below locks the goroutine to its thread around the call + check pair.
func check(rc C.int) error {
	if rc == C.PG_OK {
		return nil
	}
	return &Error{Code: int(rc), Msg: C.GoString(C.pg_last_error())}
}

func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	return check(f())
}

// Init binds the process to one GPU (pg_init).
func Init(device int) error { return call(func() C.int { return C.pg_init(C.int(device)) }) }

// InitDevices binds this ONE process to several GPUs (pg_init_devices): one context per device and one NCCL
// communicator over them, device i = rank i.  Every later call works on the device the calling OS THREAD selected
// with UseDevice, so each device is driven by a goroutine locked to its thread (compute.deviceWorker).
func InitDevices(devices []int) error {
	n := len(devices)
	if n == 0 {
		return &Error{Code: -1, Msg: "InitDevices: no devices"}
	}
	cd := (*C.int)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.int(0)))))
	defer C.free(unsafe.Pointer(cd))
	for i, d := range devices {
		*(*C.int)(unsafe.Add(unsafe.Pointer(cd), uintptr(i)*unsafe.Sizeof(C.int(0)))) = C.int(d)
	}
	return call(func() C.int { return C.pg_init_devices(C.int(n), cd) })
}

// UseDevice selects the device for the calling OS thread (the caller must hold runtime.LockOSThread for as long as
// it works on that device).  Not wrapped in call(): call's own Lock/Unlock pair nests inside the caller's lock.
func UseDevice(index int) error { return check(C.pg_use_device(C.int(index))) }

// NumDevices: devices bound by InitDevices (1 after Init).
func NumDevices() int { return int(C.pg_num_devices()) }

// Shutdown releases streams and pinned staging.
func Shutdown() error { return call(func() C.int { return C.pg_shutdown() }) }

// CommUniqueID / CommInit: multi-GPU, one process per GPU (INTEGRATION.md section 6).
func CommUniqueID() ([128]byte, error) {
	var id [128]byte
	err := call(func() C.int { return C.pg_comm_unique_id(unsafe.Pointer(&id[0])) })
	return id, err
}
func CommInit(world, rank int, id [128]byte) error {
	return call(func() C.int { return C.pg_comm_init(C.int(world), C.int(rank), unsafe.Pointer(&id[0])) })
}

// ColDesc describes one device column of a table.
type ColDesc struct {
	Name         string
	Type         int32
	Width, Scale int32
	Dict         []string // TDict8 only
}

// Table is a device-resident columnar copy of a scan's output.
type Table struct {
	h     *C.pg_table
	ncol  int
	ptrs  *C.pg_colbuf // C-allocated [ncol] scratch for Append (no Go pointer table crosses)
	Names []string
}

// NewTable creates the table; every string is copied into C memory for the call only.
func NewTable(name string, cols []ColDesc) (*Table, error) {
	n := len(cols)
	cdesc := (*C.pg_coldesc)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(C.pg_coldesc{}))))
	defer C.free(unsafe.Pointer(cdesc))
	descs := unsafe.Slice(cdesc, n)
	var frees []unsafe.Pointer
	defer func() {
		for _, p := range frees {
			C.free(p)
		}
	}()
	t := &Table{ncol: n}
	for i, c := range cols {
		cn := C.CString(c.Name)
		frees = append(frees, unsafe.Pointer(cn))
		descs[i].name = cn
		descs[i]._type = C.int32_t(c.Type)
		descs[i].width = C.int32_t(c.Width)
		descs[i].scale = C.int32_t(c.Scale)
		if len(c.Dict) > 0 {
			tab := (**C.char)(C.calloc(C.size_t(len(c.Dict)), C.size_t(unsafe.Sizeof(uintptr(0)))))
			frees = append(frees, unsafe.Pointer(tab))
			ents := unsafe.Slice(tab, len(c.Dict))
			for k, s := range c.Dict {
				ents[k] = C.CString(s)
				frees = append(frees, unsafe.Pointer(ents[k]))
			}
			descs[i].dict = tab
			descs[i].dict_len = C.int32_t(len(c.Dict))
		}
		t.Names = append(t.Names, c.Name)
	}
	cname := C.CString(name)
	defer C.free(unsafe.Pointer(cname))
	if err := call(func() C.int { return C.pg_table_create(cname, C.int(n), cdesc, &t.h) }); err != nil {
		return nil, err
	}
	t.ptrs = (*C.pg_colbuf)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(C.pg_colbuf{}))))
	return t, nil
}

// ColBuf is one flattened column of a chunk: Width bytes per value (0 = the column's native width) with a
// frame of reference, value = Base + Data[i] (pg_colbuf).  Valid is the packed validity bitmap or nil.
type ColBuf struct {
	Data  unsafe.Pointer // first byte of a Go (or C) buffer of nrows*Width bytes
	Width int32
	Base  int64
	Valid []byte
}

// Append copies nrows rows into the library's pinned staging (pg_table_append_cols).  The Go buffers are pinned
// for the call; the pointer table is C memory (t.ptrs), so cgocheck sees no Go pointer to Go pointers.
func (t *Table) Append(nrows int, cols []ColBuf) error {
	if len(cols) != t.ncol {
		return &Error{Code: EInval, Msg: "Append: column count"}
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	bufs := unsafe.Slice(t.ptrs, t.ncol)
	for i := range cols {
		pin.Pin(cols[i].Data)
		bufs[i].data = cols[i].Data
		bufs[i].width = C.int32_t(cols[i].Width)
		bufs[i].reserved = 0
		bufs[i].base = C.int64_t(cols[i].Base)
		bufs[i].valid = nil
		if len(cols[i].Valid) > 0 {
			pin.Pin(&cols[i].Valid[0])
			bufs[i].valid = (*C.uint8_t)(unsafe.Pointer(&cols[i].Valid[0]))
		}
	}
	return call(func() C.int { return C.pg_table_append_cols(t.h, C.int64_t(nrows), t.ptrs) })
}

func (t *Table) Reserve(nrows int64) error {
	return call(func() C.int { return C.pg_table_reserve(t.h, C.int64_t(nrows)) })
}
func (t *Table) Seal(globalRowOffset int64) error {
	return call(func() C.int { return C.pg_table_seal(t.h, C.int64_t(globalRowOffset)) })
}
func (t *Table) SetReplicated(on bool) error {
	d := C.int(C.PG_DIST_SHARDED)
	if on {
		d = C.int(C.PG_DIST_REPLICATED)
	}
	return call(func() C.int { return C.pg_table_set_distribution(t.h, d) })
}
func (t *Table) Rows() (int64, error) {
	var n C.int64_t
	err := call(func() C.int { return C.pg_table_rows(t.h, &n) })
	return int64(n), err
}
func (t *Table) Free() {
	if t.h != nil {
		C.pg_table_free(t.h)
		C.free(unsafe.Pointer(t.ptrs))
		t.h, t.ptrs = nil, nil
	}
}

// Plan is a compiled descriptor (include/plangpu_desc.h).
type Plan struct{ h *C.pg_plan }

func Compile(desc []int64) (*Plan, error) {
	p := &Plan{}
	var pin runtime.Pinner
	pin.Pin(&desc[0])
	defer pin.Unpin()
	err := call(func() C.int {
		return C.pg_plan_compile((*C.int64_t)(unsafe.Pointer(&desc[0])), C.size_t(len(desc)), &p.h)
	})
	if err != nil {
		return nil, err
	}
	return p, nil
}
func (p *Plan) Bind(slot int, t *Table) error {
	return call(func() C.int { return C.pg_plan_bind(p.h, C.int(slot), t.h) })
}
func (p *Plan) Prepare() error { return call(func() C.int { return C.pg_plan_prepare(p.h) }) }
func (p *Plan) Explain() string { return C.GoString(C.pg_plan_explain(p.h)) }
func (p *Plan) Execute() (*Result, error) {
	r := &Result{}
	if err := call(func() C.int { return C.pg_plan_execute(p.h, &r.h) }); err != nil {
		return nil, err
	}
	var nc C.int
	if err := call(func() C.int { return C.pg_result_num_columns(r.h, &nc) }); err != nil {
		r.Free()
		return nil, err
	}
	r.ncol = int(nc)
	r.cols = (*unsafe.Pointer)(C.calloc(C.size_t(r.ncol+1), C.size_t(unsafe.Sizeof(uintptr(0)))))
	r.valid = (**C.uint8_t)(C.calloc(C.size_t(r.ncol+1), C.size_t(unsafe.Sizeof(uintptr(0)))))
	return r, nil
}
func (p *Plan) Free() {
	if p.h != nil {
		C.pg_plan_free(p.h)
		p.h = nil
	}
}

// Result hands out <= max rows per Next; the memory belongs to the library until Free.
type Result struct {
	h     *C.pg_result
	ncol  int
	cols  *unsafe.Pointer // C-allocated [ncol]
	valid **C.uint8_t     // C-allocated [ncol]
}

func (r *Result) NumColumns() int { return r.ncol }
func (r *Result) ColumnType(col int) (typ, width, scale int32) {
	var t, w, s C.int32_t
	C.pg_result_column_type(r.h, C.int(col), &t, &w, &s)
	return int32(t), int32(w), int32(s)
}

// Next returns the row count (0 = Done) and per column a pointer to library-owned memory in the native
// encoding plus the packed validity bitmap (nil = no NULLs in this batch).
func (r *Result) Next(max int) (n int, cols []unsafe.Pointer, valid []*byte, err error) {
	var cn C.int64_t
	err = call(func() C.int {
		return C.pg_result_next(r.h, C.int64_t(max), &cn, (*unsafe.Pointer)(unsafe.Pointer(r.cols)), (**C.uint8_t)(unsafe.Pointer(r.valid)))
	})
	if err != nil || cn == 0 {
		return 0, nil, nil, err
	}
	cols = make([]unsafe.Pointer, r.ncol)
	valid = make([]*byte, r.ncol)
	copy(cols, unsafe.Slice(r.cols, r.ncol))
	for i, v := range unsafe.Slice(r.valid, r.ncol) {
		valid[i] = (*byte)(unsafe.Pointer(v))
	}
	return int(cn), cols, valid, nil
}

// Dict returns the dictionary of a TDict8 result column (nil for other columns): results are self-describing.
func (r *Result) Dict(col int) []string {
	var n C.int32_t
	var ents **C.char
	if C.pg_result_column_dict(r.h, C.int(col), &n, (***C.char)(unsafe.Pointer(&ents))) != C.PG_OK || n == 0 {
		return nil
	}
	out := make([]string, int(n))
	for i, p := range unsafe.Slice(ents, int(n)) {
		out[i] = C.GoString(p)
	}
	return out
}

func (r *Result) Free() {
	if r.h != nil {
		C.pg_result_free(r.h)
		C.free(unsafe.Pointer(r.cols))
		C.free(unsafe.Pointer(r.valid))
		r.h = nil
	}
}

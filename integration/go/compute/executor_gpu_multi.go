// executor_gpu_multi.go -- ONE plandb process driving SEVERAL GPUs (include/plangpu.h: pg_init_devices /
// pg_use_device).  The reference executes a query on one goroutine of one process (Runner.Execute, executor.go:242-296);
// here the off-loaded subtree runs on every device at once, each device on its own goroutine locked to an OS thread
// (the library keeps its device context per thread, and the NCCL collectives inside a plan -- partial-aggregate merges,
// the all-to-all row exchange -- need every rank inside pg_plan_execute concurrently).
//
// Data placement: a scanned table is drained ONCE; its 2048-row chunks are dealt out as N contiguous row ranges
// (device k seals its range with the range's global row offset, so order-dependent decimal rounding sees storage
// order); tables below cfg.Gpu.ReplicateBelow rows are appended whole to every device and marked REPLICATED.  The
// library proves co-partitioning of sharded join sides from the exchanged key ranges and otherwise ships both sides
// through its row exchange or refuses at Prepare (-> stock executors).
//
// Python twin exercised by this repository's tests: tests/multidev_check.py (threads instead of goroutines).
// Not compiled here (no Go toolchain in the image).
package compute

import (
	"runtime"
	"sync"

	"github.com/daviszhen/plan/pkg/chunk"
	"github.com/daviszhen/plan/pkg/storage"
	"github.com/daviszhen/plan/pkg/util"

	"github.com/daviszhen/plan/pkg/compute/plangpu"
)

// deviceWorker owns one device: a goroutine pinned to an OS thread that has selected the device once and then runs
// every closure sent to it.  All plangpu calls for that device go through its worker.
type deviceWorker struct {
	index int
	jobs  chan func()
}

var workers struct {
	sync.Once
	w   []*deviceWorker
	err error
}

// gpuWorkers starts the workers on first use: InitDevices on the first worker's thread, UseDevice on each.
func gpuWorkers(cfg *util.Config) ([]*deviceWorker, error) {
	workers.Do(func() {
		devs := cfg.Gpu.Devices // e.g. [0,1,2,3,4,5,6,7]
		if workers.err = plangpu.InitDevices(devs); workers.err != nil {
			return
		}
		ready := make(chan error, len(devs))
		for i := range devs {
			w := &deviceWorker{index: i, jobs: make(chan func(), 16)}
			workers.w = append(workers.w, w)
			go func() {
				runtime.LockOSThread() // never unlocked: the thread IS the device context
				ready <- plangpu.UseDevice(w.index)
				for job := range w.jobs {
					job()
				}
			}()
		}
		for range devs {
			if err := <-ready; err != nil && workers.err == nil {
				workers.err = err
			}
		}
	})
	return workers.w, workers.err
}

// onAll runs f(device index) on every worker concurrently and returns the first error.
func onAll(ws []*deviceWorker, f func(dev int) error) error {
	errs := make([]error, len(ws))
	var wg sync.WaitGroup
	for _, w := range ws {
		w := w
		wg.Add(1)
		w.jobs <- func() { defer wg.Done(); errs[w.index] = f(w.index) }
	}
	wg.Wait()
	for _, e := range errs {
		if e != nil {
			return e
		}
	}
	return nil
}

// shardedTable is the per-device copies of one scanned table.
type shardedTable struct {
	tabs       []*plangpu.Table // index = device
	version    uint64
	replicated bool
}

var shardedTables = struct {
	sync.Mutex
	m map[string]*shardedTable
}{m: map[string]*shardedTable{}}

// shardedTableFor drains the scan once (drainScan: the flattening loop of ingest in executor_gpu.go, returning the
// column descriptors and the flattened 2048-row chunkStages instead of appending them) and places the chunks.
func shardedTableFor(ws []*deviceWorker, scanOp *PhysicalOperator, cfg *util.Config, txn *storage.Txn) (*shardedTable, error) {
	si := scanOp.Info.(*ScanOpInfo)
	key := si.Database + "." + si.Table
	ver := si.TableEnt.GetStorage().Version()
	shardedTables.Lock()
	defer shardedTables.Unlock()
	if st, ok := shardedTables.m[key]; ok {
		if st.version == ver {
			return st, nil
		}
		_ = onAll(ws, func(dev int) error { st.tabs[dev].Free(); return nil })
		delete(shardedTables.m, key)
	}
	descs, stages, err := drainScan(scanOp, cfg, txn)
	if err != nil {
		return nil, err
	}
	total := int64(0)
	starts := make([]int64, len(stages)+1) // global row offset of every chunk
	for i, s := range stages {
		starts[i] = total
		total += int64(s.n)
	}
	starts[len(stages)] = total
	n := len(ws)
	st := &shardedTable{tabs: make([]*plangpu.Table, n), version: ver, replicated: total < int64(cfg.Gpu.ReplicateBelow)}
	// device k takes the chunks whose first row lies in [k*total/n, (k+1)*total/n): contiguous, in storage order
	first := make([]int, n+1)
	for k, c := 0, 0; k <= n; k++ {
		for c < len(stages) && starts[c] < int64(k)*total/int64(n) {
			c++
		}
		first[k] = c
	}
	first[n] = len(stages)
	err = onAll(ws, func(dev int) error {
		tab, err := plangpu.NewTable(key, descs)
		if err != nil {
			return err
		}
		lo, hi := first[dev], first[dev+1]
		if st.replicated {
			lo, hi = 0, len(stages)
		}
		for _, s := range stages[lo:hi] {
			if err = tab.Append(s.n, s.cols); err != nil {
				tab.Free()
				return err
			}
		}
		off := int64(0)
		if !st.replicated {
			off = starts[lo]
		}
		if err = tab.Seal(off); err != nil {
			tab.Free()
			return err
		}
		if st.replicated {
			if err = tab.SetReplicated(true); err != nil {
				tab.Free()
				return err
			}
		}
		st.tabs[dev] = tab
		return nil
	})
	if err != nil {
		return nil, err
	}
	shardedTables.m[key] = st
	return st, nil
}

// multiGpuPipelineExec is gpuPipelineExec over every device of the process.
type multiGpuPipelineExec struct {
	op    *PhysicalOperator
	cfg   *util.Config
	txn   *storage.Txn
	ws    []*deviceWorker
	plans []*plangpu.Plan
	res   []*plangpu.Result
	cur   int  // device whose result is being emitted
	rows  bool // row-emitting root over a SHARDED probe table: every device returns the rows of its shard (concatenated);
	// aggregates (merged result on every device) and row pipelines over a replicated probe table: device 0 is emitted
}

func newMultiGpuPipelineExec(op *PhysicalOperator, cfg *util.Config, txn *storage.Txn) (*multiGpuPipelineExec, error) {
	if _, err := newGpuPipelineExec(op, cfg, txn); err != nil { // same root test
		return nil, err
	}
	ws, err := gpuWorkers(cfg)
	if err != nil {
		return nil, err
	}
	rows := op.Typ == POT_Project || op.Typ == POT_Filter || op.Typ == POT_Join || op.Typ == POT_Scan
	return &multiGpuPipelineExec{op: op, cfg: cfg, txn: txn, ws: ws, rows: rows}, nil
}

func (e *multiGpuPipelineExec) Init() error {
	desc, scans, err := serializePlan(e.op)
	if err != nil {
		return err
	}
	tabs := make([]*shardedTable, len(scans))
	for slot, scanOp := range scans {
		if tabs[slot], err = shardedTableFor(e.ws, scanOp, e.cfg, e.txn); err != nil {
			return err
		}
	}
	// slot 0 is the leftmost (probe / scanned) table: serializePlan visits Children[0] first.  When it is replicated every
	// device computes the same rows, so only one copy is emitted.
	if e.rows && len(tabs) > 0 && tabs[0].replicated {
		e.rows = false
	}
	e.plans = make([]*plangpu.Plan, len(e.ws))
	// compile + bind + prepare on every device at once: Prepare agrees the statistics of sharded tables with one
	// all-gather per table, so it is a collective
	return onAll(e.ws, func(dev int) error {
		p, err := plangpu.Compile(desc)
		if err != nil {
			return err
		}
		e.plans[dev] = p
		for slot := range scans {
			if err = p.Bind(slot, tabs[slot].tabs[dev]); err != nil {
				return err
			}
		}
		return p.Prepare()
	})
}

func (e *multiGpuPipelineExec) Execute(input, output *chunk.Chunk) (OperatorResult, error) {
	ensureOutputChunk(e.op, output)
	if e.res == nil {
		e.res = make([]*plangpu.Result, len(e.ws))
		if err := onAll(e.ws, func(dev int) (err error) { e.res[dev], err = e.plans[dev].Execute(); return }); err != nil {
			return InvalidOpResult, err
		}
	}
	for e.cur < len(e.res) {
		var n int
		var err error
		dev := e.cur
		// the result lives in host memory owned by the library: reading it needs no device context
		n, cols, valid, err := e.res[dev].Next(util.DefaultVectorSize)
		if err != nil {
			return InvalidOpResult, err
		}
		if n == 0 {
			if !e.rows {
				return Done, nil // merged aggregate: identical on every device
			}
			e.cur++
			continue
		}
		for i, out := range e.op.Outputs {
			typ, _, scale := e.res[dev].ColumnType(i)
			if err = fillVector(output.Data[i], out.DataTyp, typ, scale, cols[i], valid[i], n, e.res[dev], i); err != nil {
				return InvalidOpResult, err
			}
		}
		output.SetCard(n)
		return haveMoreOutput, nil
	}
	return Done, nil
}

func (e *multiGpuPipelineExec) Close() error {
	return onAll(e.ws, func(dev int) error {
		if e.res != nil && e.res[dev] != nil {
			e.res[dev].Free()
		}
		if e.plans != nil && e.plans[dev] != nil {
			e.plans[dev].Free()
		}
		return nil
	})
}

// Package idw: GPU back end of FullSolve / ChunkSolve (drop this file and idw_entry_gpu.go next to idw.go in
// github.com/dewberry/v2r/features/idw, apply go/patches/*.patch and build with `-tags v2rgpu`).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no Go toolchain here (DESIGN.md section 2).  The
// file is the reference-side binding a v2r maintainer adds; every C call below is exercised from
// Python/ctypes and from the C++ host through exactly the same exported symbols in tests/ (see INTEGRATION.md).
//
// The map of de-duplicated points is flattened once into C-allocated pinned SoA buffers, ONE streaming
// session runs the whole exponent sweep on the GPU(s), and the unchanged processing.WriteGDAL writes
// each returned row band with the chunk-write protocol: windows of chunkR x chunkC cells at
// offsets (row0, c0), exactly like features/idw/idw_chunk.go:69-90 does per chunk.

//go:build v2rgpu

package idw

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/v2r_b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/v2r_b200/lib -lv2r_idw -Wl,-rpath,${SRCDIR}/../../third_party/v2r_b200/lib
#include <stdlib.h>
#include "v2r_idw.h"
*/
import "C"

import (
	"fmt"
	"strings"
	"sync"
	"time"
	"unsafe"

	"github.com/dewberry/v2r/tools"
	"github.com/dewberry/v2r/tools/processing"

	bunyan "github.com/Dewberry/paul-bunyan"
)

// GPUConfig is additive: the existing flags keep their meaning.
type GPUConfig struct {
	NumGPUs int  // 1, 2, 4 or 8 (row bands over NVLink); default 1
	FP32    bool // optional fast path, ~1e-5 relative
	ChunkR  int  // rows per streamed band = rows of a write window (0 = library default)
	ChunkC  int  // columns of a write window (0 = full width)
}

// GPU is the process-wide configuration the CLI flags --gpus / --fp32 fill in (cmd/idw.go).
var GPU GPUConfig

const maxSweep = 4096 // exponents per library call (include/v2r_idw.h)

var (
	gpuOnce sync.Once
	gpuCtx  *C.v2r_idw_ctx
	gpuErr  error
	gpuMu   sync.Mutex // one sweep at a time per context (the library serialises too)
)

// ctxError reads the message kept WITH the context: goroutines may migrate between OS threads from one cgo
// call to the next, so the library's thread-local v2r_idw_last_error() is only used when there is no context.
func ctxError(rc C.int) error {
	if gpuCtx != nil {
		return fmt.Errorf("v2r_idw: error %d: %s", int(rc), C.GoString(C.v2r_idw_ctx_last_error(gpuCtx)))
	}
	return fmt.Errorf("v2r_idw: error %d: %s", int(rc), C.GoString(C.v2r_idw_last_error()))
}

func gpuInit(n int) error {
	gpuOnce.Do(func() {
		if n == 0 {
			n = 1
		}
		if rc := C.v2r_idw_init(C.int(n), &gpuCtx); rc != C.V2R_IDW_OK {
			gpuErr = ctxError(rc)
		}
	})
	return gpuErr
}

// pinnedPoints flattens map[OrderedPair]Point into page-locked SoA + key arrays owned by C.
type pinnedPoints struct {
	x, y, z *C.double
	kr, kc  *C.int64_t
	n       int
}

func (p *pinnedPoints) free() {
	for _, q := range []unsafe.Pointer{unsafe.Pointer(p.x), unsafe.Pointer(p.y), unsafe.Pointer(p.z), unsafe.Pointer(p.kr), unsafe.Pointer(p.kc)} {
		C.v2r_idw_free_pinned(q)
	}
}

func flatten(data *map[tools.OrderedPair]tools.Point) (*pinnedPoints, error) {
	n := len(*data)
	p := &pinnedPoints{n: n}
	ptrs := []*unsafe.Pointer{(*unsafe.Pointer)(unsafe.Pointer(&p.x)), (*unsafe.Pointer)(unsafe.Pointer(&p.y)),
		(*unsafe.Pointer)(unsafe.Pointer(&p.z)), (*unsafe.Pointer)(unsafe.Pointer(&p.kr)), (*unsafe.Pointer)(unsafe.Pointer(&p.kc))}
	for _, q := range ptrs {
		if rc := C.v2r_idw_alloc_pinned(C.size_t(8*n), q); rc != C.V2R_IDW_OK {
			p.free()
			return nil, ctxError(rc)
		}
	}
	if n == 0 {
		return p, nil
	}
	xs := unsafe.Slice((*float64)(unsafe.Pointer(p.x)), n)
	ys := unsafe.Slice((*float64)(unsafe.Pointer(p.y)), n)
	zs := unsafe.Slice((*float64)(unsafe.Pointer(p.z)), n)
	krs := unsafe.Slice((*int64)(unsafe.Pointer(p.kr)), n)
	kcs := unsafe.Slice((*int64)(unsafe.Pointer(p.kc)), n)
	i := 0
	for key, pt := range *data { // map order is irrelevant: the key travels with the point
		xs[i], ys[i], zs[i] = pt.X, pt.Y, pt.Weight
		krs[i], kcs[i] = int64(key.R), int64(key.C)
		i++
	}
	return p, nil
}

func checkOutfile(f string) error { // same suffix rule as features/idw/idw.go:19-25 (xlsx is not a GPU output)
	if !(strings.HasSuffix(f, ".tiff") || strings.HasSuffix(f, ".tif") || strings.HasSuffix(f, ".asc")) {
		return fmt.Errorf("outfile is not supported. Use one of these files: tiff (.tif/.tiff), ascii (.asc), or excel (.xlsx).")
	}
	return nil
}

// sweepGPU is what the exponent loop of cmd/idw.go:153-161 becomes: all exponents in one pass over the points.
// outfiles[e] receives the raster of pows[e]; one completion string per exponent is sent on channel, as
// FullSolve/ChunkSolve do (features/idw/idw.go:63, features/idw/idw_chunk.go:103) -- only on success.
func sweepGPU(data *map[tools.OrderedPair]tools.Point, outfiles []string, xInfo tools.Info, yInfo tools.Info,
	proj string, pows []float64, channel chan string, cfg GPUConfig) error {
	start := time.Now()
	if len(pows) == 0 || len(pows) != len(outfiles) {
		return fmt.Errorf("one outfile per exponent is required")
	}
	for _, f := range outfiles {
		if err := checkOutfile(f); err != nil {
			return err
		}
	}
	if err := gpuInit(cfg.NumGPUs); err != nil {
		return err
	}
	numRows, numCols := tools.GetDimensions(xInfo, yInfo)
	pts, err := flatten(data)
	if err != nil {
		return err
	}
	defer pts.free()

	// a sweep longer than the library's per-call limit runs as several sessions
	for e0 := 0; e0 < len(pows); e0 += maxSweep {
		e1 := tools.Min(e0+maxSweep, len(pows))
		if err := sweepSession(pts, outfiles[e0:e1], xInfo, yInfo, numRows, numCols, proj, pows[e0:e1], cfg); err != nil {
			return err
		}
	}
	for _, pow := range pows {
		channel <- fmt.Sprintf("pow%v [%vX%v] completed in %v", pow, numRows, numCols, time.Since(start))
	}
	return nil
}

func sweepSession(pts *pinnedPoints, outfiles []string, xInfo tools.Info, yInfo tools.Info, numRows int, numCols int,
	proj string, pows []float64, cfg GPUConfig) error {
	grid := C.v2r_idw_grid{x_min: C.double(xInfo.Min), x_step: C.double(xInfo.Step), y_min: C.double(yInfo.Min),
		y_step: C.double(yInfo.Step), rows: C.int64_t(numRows), cols: C.int64_t(numCols)}
	prec := C.int(C.V2R_IDW_FP64)
	if cfg.FP32 {
		prec = C.int(C.V2R_IDW_FP32)
	}
	band := cfg.ChunkR
	if band <= 0 {
		band = numRows
		if cap := (1 << 27) / (numCols*len(pows) + 1); band > cap && cap > 0 { // ~1 GiB of float64 per band
			band = cap
		}
	}
	chunkC := cfg.ChunkC
	if chunkC <= 0 || chunkC > numCols {
		chunkC = numCols
	}
	var out unsafe.Pointer
	if rc := C.v2r_idw_alloc_pinned(C.size_t(8*band*numCols*len(pows)), &out); rc != C.V2R_IDW_OK {
		return ctxError(rc)
	}
	defer C.v2r_idw_free_pinned(out)

	gpuMu.Lock()
	defer gpuMu.Unlock()
	// _stream_begin returns only after the pinned point arrays have been read completely (include/v2r_idw.h)
	if rc := C.v2r_idw_stream_begin(gpuCtx, pts.x, pts.y, pts.z, pts.kr, pts.kc, C.int64_t(pts.n), &grid,
		(*C.double)(unsafe.Pointer(&pows[0])), C.int(len(pows)), prec, C.int64_t(band)); rc != C.V2R_IDW_OK {
		return ctxError(rc)
	}
	defer C.v2r_idw_stream_end(gpuCtx)

	totalSize := tools.RCToPair(numRows, numCols)
	gdal := processing.CreateGDalInfo(xInfo.Min, yInfo.Min, xInfo.Step, yInfo.Step, 7, proj)
	created := make([]bool, len(pows))
	totalBands := tools.Max(1, tools.RoundUp(numRows, band))
	progress := tools.Max(1, totalBands/10)
	var window []float64 // contiguous copy of a chunkC-wide window (WriteGDAL passes no line stride to GDAL)
	for i := 0; ; i++ {
		var row0, nrows C.int64_t
		if rc := C.v2r_idw_stream_next(gpuCtx, (*C.double)(out), &row0, &nrows); rc != C.V2R_IDW_OK {
			return ctxError(rc)
		}
		if nrows == 0 {
			break
		}
		nr := int(nrows)
		plane := nr * numCols
		all := unsafe.Slice((*float64)(out), plane*len(pows))
		for e := range pows { // the next band is already computing while these windows are written
			flat := all[e*plane : (e+1)*plane]
			for c0 := 0; c0 < numCols; c0 += chunkC {
				wc := tools.Min(chunkC, numCols-c0)
				buf := flat
				if wc != numCols {
					if cap(window) < nr*wc {
						window = make([]float64, nr*wc)
					}
					buf = window[:nr*wc]
					for r := 0; r < nr; r++ {
						copy(buf[r*wc:(r+1)*wc], flat[r*numCols+c0:r*numCols+c0+wc])
					}
				}
				if err := processing.WriteGDAL(buf, gdal, outfiles[e], tools.MakePair(int(row0), c0), totalSize,
					tools.MakePair(nr, wc), !created[e]); err != nil {
					return err
				}
				created[e] = true
			}
		}
		if cfg.ChunkR > 0 && (i+1)%progress == 0 { // features/idw/idw_chunk.go:92-94
			bunyan.Infof("~%d%%, %v / %v", 100*(i+1)/totalBands, i+1, totalBands)
		}
	}
	return nil
}

// Package idw: GPU back end for FullSolve / ChunkSolve (drop this file next to idw.go in
// github.com/dewberry/v2r/features/idw and build with `-tags v2rgpu`).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no Go toolchain here (DESIGN.md section 2).  The
// file is the reference-side binding a v2r maintainer adds; every C call below is exercised from
// Python/ctypes through exactly the same exported symbols in tests/ (see INTEGRATION.md).
//
// It keeps the reference's Go signatures (features/idw/idw.go:15, features/idw/idw_chunk.go:43):
// the map of de-duplicated points is flattened once into C-allocated pinned SoA buffers, ONE cgo
// call runs the whole exponent sweep on the GPU(s), and the unchanged processing.WriteGDAL writes
// each returned row band with the chunk-write protocol (offsets = (row0, 0)).

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
)

// GPUConfig is additive: the existing flags keep their meaning.
type GPUConfig struct {
	NumGPUs   int  // 1, 2, 4 or 8 (row bands over NVLink); default 1
	FP32      bool // optional fast path, ~1e-5 relative
	BandRows  int  // rows per streamed band (0 = library default, ~1 GiB of raster)
}

var (
	gpuOnce sync.Once
	gpuCtx  *C.v2r_idw_ctx
	gpuErr  error
	gpuMu   sync.Mutex // one sweep at a time per context (the library serialises too)
)

func lastError(rc C.int) error {
	return fmt.Errorf("v2r_idw: error %d: %s", int(rc), C.GoString(C.v2r_idw_last_error()))
}

func gpuInit(n int) error {
	gpuOnce.Do(func() {
		if n == 0 {
			n = 1
		}
		if rc := C.v2r_idw_init(C.int(n), &gpuCtx); rc != C.V2R_IDW_OK {
			gpuErr = lastError(rc)
		}
	})
	return gpuErr
}

// pinnedPoints flattens map[OrderedPair]Point into page-locked SoA + key arrays owned by C.
type pinnedPoints struct {
	x, y, z    *C.double
	kr, kc     *C.int64_t
	n          int
}

func (p *pinnedPoints) free() {
	for _, q := range []unsafe.Pointer{unsafe.Pointer(p.x), unsafe.Pointer(p.y), unsafe.Pointer(p.z), unsafe.Pointer(p.kr), unsafe.Pointer(p.kc)} {
		C.v2r_idw_free_pinned(q)
	}
}

func flatten(data *map[tools.OrderedPair]tools.Point) (*pinnedPoints, error) {
	n := len(*data)
	p := &pinnedPoints{n: n}
	alloc := func(dst *unsafe.Pointer) error {
		if rc := C.v2r_idw_alloc_pinned(C.size_t(8*n), dst); rc != C.V2R_IDW_OK {
			return lastError(rc)
		}
		return nil
	}
	ptrs := []*unsafe.Pointer{(*unsafe.Pointer)(unsafe.Pointer(&p.x)), (*unsafe.Pointer)(unsafe.Pointer(&p.y)),
		(*unsafe.Pointer)(unsafe.Pointer(&p.z)), (*unsafe.Pointer)(unsafe.Pointer(&p.kr)), (*unsafe.Pointer)(unsafe.Pointer(&p.kc))}
	for _, q := range ptrs {
		if err := alloc(q); err != nil {
			return nil, err
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

// SweepSolveGPU is what the exponent loop of cmd/idw.go:153-161 becomes: all exponents in one pass over
// the points.  outfiles[e] receives the raster of pows[e]; one completion string per exponent is sent on
// channel, as FullSolve/ChunkSolve do (features/idw/idw.go:63, features/idw/idw_chunk.go:103).
func SweepSolveGPU(data *map[tools.OrderedPair]tools.Point, outfiles []string, xInfo tools.Info, yInfo tools.Info,
	proj string, pows []float64, channel chan string, cfg GPUConfig) error {
	start := time.Now()
	for _, f := range outfiles { // same suffix rule as features/idw/idw.go:19-25 (xlsx is not a GPU output)
		if !(strings.HasSuffix(f, ".tiff") || strings.HasSuffix(f, ".tif") || strings.HasSuffix(f, ".asc")) {
			return fmt.Errorf("outfile is not supported. Use one of these files: tiff (.tif/.tiff), ascii (.asc), or excel (.xlsx).")
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

	grid := C.v2r_idw_grid{x_min: C.double(xInfo.Min), x_step: C.double(xInfo.Step), y_min: C.double(yInfo.Min),
		y_step: C.double(yInfo.Step), rows: C.int64_t(numRows), cols: C.int64_t(numCols)}
	prec := C.int(C.V2R_IDW_FP64)
	if cfg.FP32 {
		prec = C.int(C.V2R_IDW_FP32)
	}
	band := cfg.BandRows
	if band <= 0 {
		band = numRows
		if cap := (1 << 27) / (numCols*len(pows) + 1); band > cap && cap > 0 { // ~1 GiB of float64 per band
			band = cap
		}
	}
	var out unsafe.Pointer
	if rc := C.v2r_idw_alloc_pinned(C.size_t(8*band*numCols*len(pows)), &out); rc != C.V2R_IDW_OK {
		return lastError(rc)
	}
	defer C.v2r_idw_free_pinned(out)

	gpuMu.Lock()
	defer gpuMu.Unlock()
	if rc := C.v2r_idw_stream_begin(gpuCtx, pts.x, pts.y, pts.z, pts.kr, pts.kc, C.int64_t(pts.n), &grid,
		(*C.double)(unsafe.Pointer(&pows[0])), C.int(len(pows)), prec, C.int64_t(band)); rc != C.V2R_IDW_OK {
		return lastError(rc)
	}
	defer C.v2r_idw_stream_end(gpuCtx)

	totalSize := tools.RCToPair(numRows, numCols)
	gdal := processing.CreateGDalInfo(xInfo.Min, yInfo.Min, xInfo.Step, yInfo.Step, 7, proj)
	created := make([]bool, len(pows))
	for {
		var row0, nrows C.int64_t
		if rc := C.v2r_idw_stream_next(gpuCtx, (*C.double)(out), &row0, &nrows); rc != C.V2R_IDW_OK {
			return lastError(rc)
		}
		if nrows == 0 {
			break
		}
		plane := int(nrows) * numCols
		all := unsafe.Slice((*float64)(out), plane*len(pows))
		for e := range pows { // the next band is already computing while these windows are written
			flat := all[e*plane : (e+1)*plane]
			if err := processing.WriteGDAL(flat, gdal, outfiles[e], tools.MakePair(int(row0), 0), totalSize,
				tools.MakePair(int(nrows), numCols), !created[e]); err != nil {
				return err
			}
			created[e] = true
		}
	}
	for _, pow := range pows {
		channel <- fmt.Sprintf("pow%v [%vX%v] completed in %v", pow, numRows, numCols, time.Since(start))
	}
	return nil
}

// FullSolveGPU / ChunkSolveGPU keep the reference signatures so cmd/idw.go and idw_test.go can switch by
// renaming one call.  chunkR becomes the streamed band height; chunkC has no GPU meaning (bands are
// full-width windows, which is what WriteGDAL prefers).
func FullSolveGPU(data *map[tools.OrderedPair]tools.Point, outfile string, xInfo tools.Info, yInfo tools.Info,
	proj string, pow float64, channel chan string) error {
	return SweepSolveGPU(data, []string{outfile}, xInfo, yInfo, proj, []float64{pow}, channel, GPUConfig{})
}

func ChunkSolveGPU(data *map[tools.OrderedPair]tools.Point, outfile string, xInfo tools.Info, yInfo tools.Info,
	chunkR int, chunkC int, proj string, pow float64, channel chan string) error {
	numRows, numCols := tools.GetDimensions(xInfo, yInfo)
	if chunkR > numRows || chunkC > numCols { // features/idw/idw_chunk.go:48-53
		chunkR = numRows / 4
	}
	return SweepSolveGPU(data, []string{outfile}, xInfo, yInfo, proj, []float64{pow}, channel, GPUConfig{BandRows: chunkR})
}

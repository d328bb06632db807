// FullSolve / ChunkSolve / SweepSolve under `-tags v2rgpu`: the reference's own entry points
// (features/idw/idw.go:15, features/idw/idw_chunk.go:43), same signatures, dispatching to CUDA through
// idw_gpu.go.  go/patches/0001-build-tags.patch puts `//go:build !v2rgpu` on the two reference files that
// define the CPU versions, so exactly one definition of each name exists in either build;
// features/idw/idw_test.go and cmd/idw.go call the same names in both.

//go:build v2rgpu

package idw

import (
	"github.com/dewberry/v2r/tools"

	bunyan "github.com/Dewberry/paul-bunyan"
)

// Run IDW on the GPU(s) for the whole grid. Print to outfile
func FullSolve(data *map[tools.OrderedPair]tools.Point, outfile string, xInfo tools.Info, yInfo tools.Info, proj string, pow float64, channel chan string) error {
	cfg := GPU
	cfg.ChunkR, cfg.ChunkC = 0, 0
	return sweepGPU(data, []string{outfile}, xInfo, yInfo, proj, []float64{pow}, channel, cfg)
}

// Same values as FullSolve, streamed out of the device in chunkR-row bands and written in chunkR x chunkC windows.
func ChunkSolve(data *map[tools.OrderedPair]tools.Point, outfile string, xInfo tools.Info, yInfo tools.Info, chunkR int, chunkC int, proj string, pow float64, channel chan string) error {
	chunkR, chunkC = clampChunks(xInfo, yInfo, chunkR, chunkC)
	cfg := GPU
	cfg.ChunkR, cfg.ChunkC = chunkR, chunkC
	return sweepGPU(data, []string{outfile}, xInfo, yInfo, proj, []float64{pow}, channel, cfg)
}

// SweepSolve runs every exponent of the sweep from ONE pass over the points (cmd/idw.go's exponent loop);
// one completion message per exponent arrives on channel, so the caller's drain loop is unchanged.
func SweepSolve(data *map[tools.OrderedPair]tools.Point, outfiles []string, xInfo tools.Info, yInfo tools.Info,
	useChunking bool, chunkR int, chunkC int, proj string, pows []float64, channel chan string) {
	cfg := GPU
	cfg.ChunkR, cfg.ChunkC = 0, 0
	if useChunking {
		cfg.ChunkR, cfg.ChunkC = clampChunks(xInfo, yInfo, chunkR, chunkC)
	}
	go func() {
		if err := sweepGPU(data, outfiles, xInfo, yInfo, proj, pows, channel, cfg); err != nil {
			bunyan.Fatal(err) // the CPU build drops solver errors and then blocks forever on the channel
		}
	}()
}

func clampChunks(xInfo tools.Info, yInfo tools.Info, chunkR int, chunkC int) (int, int) {
	numRows, numCols := tools.GetDimensions(xInfo, yInfo)
	if chunkR > numRows || chunkC > numCols { // features/idw/idw_chunk.go:48-53
		bunyan.Warnf("chunk size, %v too large for total image %v", tools.MakePair(chunkR, chunkC), tools.MakePair(numRows, numCols))
		bunyan.Warn("Chunk sizes changed to ~1/16 of total size")
		chunkR = numRows / 4
		chunkC = numCols / 4
	}
	return chunkR, chunkC
}

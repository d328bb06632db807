// SweepSolve for the stock CPU build (no tag): the exponent loop of cmd/idw.go:153-161 moved behind one name, so
// cmd/idw.go is identical in both builds.  One goroutine per exponent, as before.

//go:build !v2rgpu

package idw

import "github.com/dewberry/v2r/tools"

// GPUConfig exists in both builds so that cmd/idw.go can fill it from --gpus / --fp32; the CPU build ignores it.
type GPUConfig struct {
	NumGPUs int
	FP32    bool
	ChunkR  int
	ChunkC  int
}

var GPU GPUConfig

func SweepSolve(data *map[tools.OrderedPair]tools.Point, outfiles []string, xInfo tools.Info, yInfo tools.Info,
	useChunking bool, chunkR int, chunkC int, proj string, pows []float64, channel chan string) {
	for i, pow := range pows {
		if !useChunking {
			go FullSolve(data, outfiles[i], xInfo, yInfo, proj, pow, channel)
		} else {
			go ChunkSolve(data, outfiles[i], xInfo, yInfo, chunkR, chunkC, proj, pow, channel)
		}
	}
}

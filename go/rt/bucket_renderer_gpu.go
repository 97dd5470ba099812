// bucket_renderer_gpu.go — what changes in rt/bucket_renderer.go for the B200 drop-in. Two edits to the reference file:
//
//   1. NewBucketRenderer (rt/bucket_renderer.go:54-74): after the struct is filled, call r.initGPU() (below). Signature unchanged;
//      bucketSize and numWorkers are accepted and ignored — the library fans the pass out to the GPUs itself.
//   2. renderPass (rt/bucket_renderer.go:170-214): replace the body after the `switch r.currentPass` block (the bucket channel,
//      the worker goroutines and wg.Wait()) with `r.renderPassGPU(samplesForPass, depthForPass)`.
//
// Update() keeps starting renderMultiPass in a goroutine and polling passComplete (rt/bucket_renderer.go:127-164): the ebiten
// thread never blocks on a pass, exactly as today. Draw / SaveImage / drawStatsToFramebuffer / IsCompleted are untouched — they
// read r.framebuffer under r.mu.
//
// NOT COMPILED IN THE BUILD CONTAINER (no Go toolchain): go-raytracing_b200/host/rt_renderer.cpp is the tested mirror.
package rt

import (
	"fmt"
	"time"
)

// fields added to BucketRenderer:
//     gpu      *gpuContext
//     gpuSeed  uint64

// initGPU: rtx_create_multi over every visible GPU (RT_GPUS selects), flatten + upload the scene, set the camera. Runs once, in
// NewBucketRenderer. A scene the device path cannot run (user-defined Hittable, RotateX/Z ...) panics with the flattener's
// message, as LoadOBJ failures do in the scene functions (rt/scenes.go:771-773): there is no CPU fallback to fall back to.
func (r *BucketRenderer) initGPU() {
	g, err := newGPUContext()
	if err != nil {
		panic(fmt.Sprintf("BucketRenderer: %v", err))
	}
	fs, err := flattenScene(r.world, r.camera)
	if err == nil {
		err = g.uploadScene(fs)
	}
	if err == nil {
		err = g.setCamera(r.camera)
	}
	if err != nil {
		g.Close()
		panic(fmt.Sprintf("BucketRenderer: %v", err))
	}
	r.gpu = g
	r.gpuSeed = uint64(time.Now().UnixNano()) // the reference draws from Go's auto-seeded source: every run differs
}

// renderPassGPU is the new tail of renderPass: one blocking call per pass, on the goroutine Update() started for the pass.
// The pass renders into a private buffer and is swapped into the framebuffer under the mutex, so Draw() never sees a
// half-written image (the reference writes bucket by bucket under the same mutex, rt/bucket_renderer.go:291-300).
func (r *BucketRenderer) renderPassGPU(samplesForPass, depthForPass int) {
	pix := make([]uint8, len(r.framebuffer.Pix))
	seed := r.gpuSeed + uint64(r.currentPass)*0x9E3779B97F4A7C15
	if err := r.gpu.renderPass(samplesForPass, depthForPass, r.camera.MaxDepth, seed, pix); err != nil {
		panic(fmt.Sprintf("BucketRenderer: %v", err))
	}
	r.mu.Lock()
	copy(r.framebuffer.Pix, pix)
	r.mu.Unlock()
	// the counters the progress overlay and main.go's summary read (rt/bucket_renderer.go:272, :287; rt/profiler.go)
	st := r.gpu.stats()
	r.completedCount.Store(int32(r.totalBuckets))
	GlobalRenderStats.SamplesComputed.Add(int64(st.Paths))
	GlobalRenderStats.RayCount.Add(int64(st.Paths + st.ExtensionRays)) // RayCount counts RayColor calls: rt/camera.go:439 + :448
	GlobalRenderStats.PixelsRendered.Add(int64(r.camera.ImageWidth * r.camera.ImageHeight))
	r.passComplete.Store(true)
}

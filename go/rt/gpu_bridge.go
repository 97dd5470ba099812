// gpu_bridge.go — the cgo binding of librtx_b200.so (include/rtx_b200.h) for package rt of byvfx/go-raytracing.
//
// Drop this file, flatten.go and bucket_renderer_gpu.go into the reference's rt/ directory (they use the package's unexported
// fields, like the rest of rt/). Nothing above BucketRenderer.renderPass changes: NewBucketRenderer keeps its signature
// (rt/bucket_renderer.go:54), Update / Draw / SaveImage / IsCompleted / GetRenderDuration keep their bodies, main.go keeps building
// one renderer (main.go:83-93). What changes is the body of renderPass (rt/bucket_renderer.go:170-214): instead of feeding 32x32
// buckets to numWorkers goroutines it makes ONE call per pass into the CUDA library, which fans the pass out to the GPUs.
//
// NOT COMPILED IN THE BUILD CONTAINER (no Go toolchain there, see DESIGN.md section 1): the C++ host mirror
// (go-raytracing_b200/host/rt_flatten.cpp, rt_renderer.cpp) makes the same calls in the same order and is what the tests drive.
//
// cgo pointer rules: every Go slice handed to C is only read for the duration of the call (the library copies), and no Go pointer
// is stored on the C side; the arrays of a scene description are pinned with runtime.Pinner for the one rtx_scene_upload call
// because rtx_scene_desc itself holds pointers to them.
package rt

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../go-raytracing_b200/csrc -lrtx_b200 -Wl,-rpath,${SRCDIR}/../../go-raytracing_b200/csrc
#include <stdlib.h>
#include "rtx_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"os"
	"runtime"
	"strconv"
	"strings"
	"unsafe"
)

// gpuContext owns one rtx_ctx: one GPU, or — the default on a multi-GPU box — every visible GPU behind one context
// (rtx_create_multi: the library slices the samples of a pass over the devices and sums the buffers with one ncclReduce).
type gpuContext struct {
	h       *C.rtx_ctx
	devices []int32
	width   int
	height  int
}

func (g *gpuContext) lastError() error {
	return errors.New(C.GoString(C.rtx_last_error(g.h)))
}

func check(g *gpuContext, rc C.int32_t, what string) error {
	if int32(rc) == int32(C.RTX_OK) { // (an enumerator: convert instead of relying on how cgo types it)
		return nil
	}
	var msg string
	if g != nil && g.h != nil {
		msg = C.GoString(C.rtx_last_error(g.h))
	} else {
		msg = C.GoString(C.rtx_last_error(nil))
	}
	return fmt.Errorf("%s = %d: %s", what, int(rc), msg)
}

// gpuDevices: RT_GPUS="0,1,2,3" selects devices; unset = all visible devices. There is no CPU fallback: without a device
// newGPUContext returns the library's error and NewBucketRenderer panics, exactly as a missing asset does in the reference.
func gpuDevices() ([]int32, error) {
	if env := os.Getenv("RT_GPUS"); env != "" {
		var ids []int32
		for _, tok := range strings.Split(env, ",") {
			v, err := strconv.Atoi(strings.TrimSpace(tok))
			if err != nil {
				return nil, fmt.Errorf("RT_GPUS: %v", err)
			}
			ids = append(ids, int32(v))
		}
		return ids, nil
	}
	n := int(C.rtx_device_count())
	if n == 0 {
		return []int32{0}, nil // rtx_create reports "no CUDA device ... no CPU fallback"
	}
	ids := make([]int32, n)
	for i := range ids {
		ids[i] = int32(i)
	}
	return ids, nil
}

func newGPUContext() (*gpuContext, error) {
	ids, err := gpuDevices()
	if err != nil {
		return nil, err
	}
	g := &gpuContext{devices: ids}
	rc := C.rtx_create_multi((*C.int32_t)(unsafe.Pointer(&ids[0])), C.int32_t(len(ids)), &g.h)
	if err := check(nil, rc, "rtx_create_multi"); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(g, func(g *gpuContext) { g.Close() })
	return g, nil
}

func (g *gpuContext) Close() {
	if g.h != nil {
		C.rtx_destroy(g.h)
		g.h = nil
	}
}

// cArr returns a C view of a Go slice (nil for an empty one) and pins it until the Pinner is released.
func cF64(p *runtime.Pinner, s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	p.Pin(&s[0])
	return (*C.double)(unsafe.Pointer(&s[0]))
}
func cI32(p *runtime.Pinner, s []int32) *C.int32_t {
	if len(s) == 0 {
		return nil
	}
	p.Pin(&s[0])
	return (*C.int32_t)(unsafe.Pointer(&s[0]))
}
func cI64(p *runtime.Pinner, s []int64) *C.int64_t {
	if len(s) == 0 {
		return nil
	}
	p.Pin(&s[0])
	return (*C.int64_t)(unsafe.Pointer(&s[0]))
}

// uploadScene: rtx_scene_upload of a flattened scene (flatten.go). The library copies everything and builds the wide BVHs on
// the device; the Go arrays can be garbage-collected afterwards.
func (g *gpuContext) uploadScene(fs *flatScene) error {
	var pin runtime.Pinner
	defer pin.Unpin()
	var d C.rtx_scene_desc
	d.abi_version = C.RTX_ABI_VERSION
	if fs.worldIsBVH {
		d.world_is_bvh = 1
	}
	d.n_textures = C.int32_t(len(fs.texType))
	d.tex_type, d.tex_color, d.tex_inv_scale = cI32(&pin, fs.texType), cF64(&pin, fs.texColor), cF64(&pin, fs.texInvScale)
	d.tex_even, d.tex_odd = cI32(&pin, fs.texEven), cI32(&pin, fs.texOdd)
	d.n_materials = C.int32_t(len(fs.matType))
	d.mat_type, d.mat_tex, d.mat_albedo = cI32(&pin, fs.matType), cI32(&pin, fs.matTex), cF64(&pin, fs.matAlbedo)
	d.mat_fuzz, d.mat_ior = cF64(&pin, fs.matFuzz), cF64(&pin, fs.matIor)
	d.n_spheres = C.int32_t(len(fs.sphMat))
	d.sph_center, d.sph_velocity, d.sph_radius, d.sph_mat = cF64(&pin, fs.sphCenter), cF64(&pin, fs.sphVelocity), cF64(&pin, fs.sphRadius), cI32(&pin, fs.sphMat)
	d.n_quads = C.int32_t(len(fs.quadMat))
	d.quad_q, d.quad_u, d.quad_v, d.quad_mat = cF64(&pin, fs.quadQ), cF64(&pin, fs.quadU), cF64(&pin, fs.quadV), cI32(&pin, fs.quadMat)
	d.n_tris = C.int32_t(len(fs.triMat))
	d.tri_v0, d.tri_v1, d.tri_v2 = cF64(&pin, fs.triV0), cF64(&pin, fs.triV1), cF64(&pin, fs.triV2)
	d.tri_mat, d.tri_rank = cI32(&pin, fs.triMat), cI32(&pin, fs.triRank)
	d.n_planes = C.int32_t(len(fs.planeMat))
	d.plane_point, d.plane_normal, d.plane_mat = cF64(&pin, fs.planePoint), cF64(&pin, fs.planeNormal), cI32(&pin, fs.planeMat)
	d.n_circles = C.int32_t(len(fs.circleMat))
	d.circle_center, d.circle_normal, d.circle_radius, d.circle_mat = cF64(&pin, fs.circleCenter), cF64(&pin, fs.circleNormal), cF64(&pin, fs.circleRadius), cI32(&pin, fs.circleMat)
	d.n_perlin = C.int32_t(len(fs.perlinPerm) / 768)
	d.perlin_vec, d.perlin_perm = cF64(&pin, fs.perlinVec), cI32(&pin, fs.perlinPerm)
	d.n_images = C.int32_t(len(fs.imageWidth))
	d.image_width, d.image_height, d.image_offset, d.image_rgb = cI32(&pin, fs.imageWidth), cI32(&pin, fs.imageHeight), cI64(&pin, fs.imageOffset), cF64(&pin, fs.imageRGB)
	d.n_groups = C.int32_t(len(fs.groupKind))
	d.group_kind, d.group_begin, d.group_count = cI32(&pin, fs.groupKind), cI32(&pin, fs.groupBegin), cI32(&pin, fs.groupCount)
	d.n_list_items = C.int32_t(len(fs.listItemKind))
	d.list_item_kind, d.list_item_index = cI32(&pin, fs.listItemKind), cI32(&pin, fs.listItemIndex)
	d.n_xforms = C.int32_t(len(fs.xfType))
	d.xf_type, d.xf_a, d.xf_b = cI32(&pin, fs.xfType), cF64(&pin, fs.xfA), cF64(&pin, fs.xfB)
	d.n_volumes = C.int32_t(len(fs.volMat))
	d.vol_neg_inv_density, d.vol_mat = cF64(&pin, fs.volNegInvDensity), cI32(&pin, fs.volMat)
	d.n_entries = C.int32_t(len(fs.entryKind))
	d.entry_geom_kind, d.entry_geom_index = cI32(&pin, fs.entryKind), cI32(&pin, fs.entryIndex)
	d.entry_xf_begin, d.entry_xf_count = cI32(&pin, fs.entryXfBegin), cI32(&pin, fs.entryXfCount)
	d.entry_volume, d.entry_rank = cI32(&pin, fs.entryVolume), cI32(&pin, fs.entryRank)
	d.n_lights = C.int32_t(len(fs.lightQuad))
	d.light_quad = cI32(&pin, fs.lightQuad)
	if fs.envWidth > 0 {
		d.env_width, d.env_height = C.int32_t(fs.envWidth), C.int32_t(fs.envHeight)
		d.env_rgb = cF64(&pin, fs.envRGB)
		d.env_rotation = C.double(fs.envRotation)
		if fs.envImportanceSampling {
			d.env_importance_sampling = 1
		}
	}
	return check(g, C.rtx_scene_upload(g.h, &d), "rtx_scene_upload")
}

// setCamera: rtx_camera_set with the camera's post-Initialize state passed verbatim (has_derived = 1), so that the device uses
// the very float64 values Go's math.Tan produced (rt/camera.go:286-344).
func (g *gpuContext) setCamera(c *Camera) error {
	var d C.rtx_camera_desc
	v3 := func(dst *[3]C.double, v Vec3) { dst[0], dst[1], dst[2] = C.double(v.X), C.double(v.Y), C.double(v.Z) }
	b := func(x bool) C.int32_t {
		if x {
			return 1
		}
		return 0
	}
	d.aspect_ratio, d.image_width = C.double(c.AspectRatio), C.int32_t(c.ImageWidth)
	d.samples_per_pixel, d.max_depth, d.vfov = C.int32_t(c.SamplesPerPixel), C.int32_t(c.MaxDepth), C.double(c.Vfov)
	v3(&d.look_from, c.LookFrom)
	v3(&d.look_at, c.LookAt)
	v3(&d.vup, c.Vup)
	d.defocus_angle, d.focus_dist = C.double(c.DefocusAngle), C.double(c.FocusDist)
	v3(&d.look_from2, c.LookFrom2)
	v3(&d.look_at2, c.LookAt2)
	d.camera_motion, d.free_camera = b(c.CameraMotion), b(c.FreeCamera)
	v3(&d.forward, c.Forward)
	v3(&d.background, c.Background)
	d.use_sky_gradient, d.phantom_hdri = b(c.UseSkyGradient), b(c.PhantomHDRI)
	d.has_derived, d.image_height = 1, C.int32_t(c.ImageHeight)
	v3(&d.center, c.center)
	v3(&d.pixel00_loc, c.pixel00Loc)
	v3(&d.pixel_delta_u, c.pixelDeltaU)
	v3(&d.pixel_delta_v, c.pixelDeltaV)
	v3(&d.u, c.u)
	v3(&d.v, c.v)
	v3(&d.w, c.w)
	// rt/camera.go:356: FocusDist * tan(radians(DefocusAngle / 2)), recomputed per ray by the reference
	d.defocus_radius = C.double(c.defocusDiskU.Len())
	d.viewport_width, d.viewport_height = C.double(c.viewportWidth), C.double(c.viewportHeight)
	if err := check(g, C.rtx_camera_set(g.h, &d), "rtx_camera_set"); err != nil {
		return err
	}
	g.width, g.height = c.ImageWidth, c.ImageHeight
	return nil
}

// renderPass: one pass of `spp` samples per pixel at depth `depth` into a cleared accumulation buffer, resolved into pix
// (framebuffer.Pix: row-major RGBA8, stride 4*W, A = 255 — image.NewRGBA, rt/bucket_renderer.go:55). Blocking: call it from the
// goroutine that used to run renderPass.
func (g *gpuContext) renderPass(spp, depth, cameraMaxDepth int, seed uint64, pix []uint8) error {
	if err := check(g, C.rtx_accum_clear(g.h), "rtx_accum_clear"); err != nil {
		return err
	}
	rc := C.rtx_render_pass(g.h, C.int32_t(spp), C.int32_t(depth), C.int32_t(cameraMaxDepth), C.uint64_t(seed), 0)
	if err := check(g, rc, "rtx_render_pass"); err != nil {
		return err
	}
	rc = C.rtx_resolve_rgba8(g.h, C.int32_t(spp), (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.int64_t(len(pix)))
	return check(g, rc, "rtx_resolve_rgba8")
}

// gpuStats mirrors the fields of rtx_stats the renderer prints (the reference prints its own counters in main.go:95-105).
type gpuStats struct {
	Paths, ExtensionRays, ShadowRays uint64
	MsTotal, MsReduce, MsTail        float64
	Devices                          int
}

func (g *gpuContext) stats() gpuStats {
	var s C.rtx_stats
	C.rtx_get_stats(g.h, &s)
	return gpuStats{uint64(s.paths), uint64(s.extension_rays), uint64(s.shadow_rays), float64(s.ms_total), float64(s.ms_reduce), float64(s.ms_tail), int(s.n_devices)}
}

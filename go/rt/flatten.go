// flatten.go — the reference's pointer scene graph -> the SoA arrays of rtx_scene_desc (include/rtx_b200.h).
//
// Lives in package rt because it reads unexported fields (Quad.u/v/mat, Triangle.v0..v2, BVHNode.left/right, Lambertian.tex,
// Volume.boundary ...). It is a type switch over the concrete types of rt/; a user-defined Hittable / Material / Texture is a
// flatten-time error — never a silent CPU fallback. Mirrors go-raytracing_b200/host/rt_flatten.cpp line for line (that file is
// what the tests exercise; this one cannot be compiled in the build container).
package rt

import "fmt"

// enum values of include/rtx_b200.h
const (
	rtxMatLambertian, rtxMatMetal, rtxMatDielectric, rtxMatDiffuseLight, rtxMatIsotropic = 0, 1, 2, 3, 4
	rtxTexSolid, rtxTexChecker, rtxTexNoise, rtxTexImage                              = 0, 1, 2, 3
	rtxGeomSphere, rtxGeomQuad, rtxGeomTriangle, rtxGeomPlane, rtxGeomList, rtxGeomMesh, rtxGeomCircle = 0, 1, 2, 3, 4, 5, 6
	rtxXfTranslate, rtxXfRotateY, rtxXfScale                                          = 0, 1, 2
)

type flatScene struct {
	worldIsBVH bool

	texType, texEven, texOdd []int32
	texColor, texInvScale    []float64

	matType, matTex             []int32
	matAlbedo, matFuzz, matIor  []float64

	sphCenter, sphVelocity, sphRadius []float64
	sphMat                            []int32
	quadQ, quadU, quadV               []float64
	quadMat                           []int32
	triV0, triV1, triV2               []float64
	triMat, triRank                   []int32
	planePoint, planeNormal           []float64
	planeMat                          []int32
	circleCenter, circleNormal, circleRadius []float64
	circleMat                         []int32

	perlinVec  []float64
	perlinPerm []int32

	imageWidth, imageHeight []int32
	imageOffset             []int64
	imageRGB                []float64

	groupKind, groupBegin, groupCount []int32
	listItemKind, listItemIndex       []int32

	xfType   []int32
	xfA, xfB []float64

	volNegInvDensity []float64
	volMat           []int32

	entryKind, entryIndex, entryXfBegin, entryXfCount, entryVolume, entryRank []int32

	lightQuad []int32

	envWidth, envHeight   int
	envRGB                []float64
	envRotation           float64
	envImportanceSampling bool
}

type flattener struct {
	fs       *flatScene
	texIDs   map[Texture]int32
	matIDs   map[Material]int32
	primIDs  map[Hittable]int32 // spheres, quads, triangles, planes, circles: shared objects are stored once
	groupIDs map[Hittable]int32
	perlins  map[*Perlin]int32
}

func push3(dst *[]float64, v Vec3) { *dst = append(*dst, v.X, v.Y, v.Z) }

func (f *flattener) texture(t Texture) (int32, error) {
	if t == nil {
		return 0, fmt.Errorf("flatten: nil texture")
	}
	if id, ok := f.texIDs[t]; ok {
		return id, nil
	}
	fs := f.fs
	typ, even, odd := int32(rtxTexSolid), int32(-1), int32(-1)
	var color Color
	inv := 0.0
	switch x := t.(type) {
	case *SolidColor:
		color = x.Albedo
	case *CheckerTexture:
		typ, inv = rtxTexChecker, x.invScale
		var err error
		if even, err = f.texture(x.even); err != nil {
			return 0, err
		}
		if odd, err = f.texture(x.odd); err != nil {
			return 0, err
		}
	case *NoiseTexture: // scale travels in tex_inv_scale (not inverted), the Perlin table index in tex_even
		typ, inv = rtxTexNoise, x.scale
		id, ok := f.perlins[x.noise]
		if !ok {
			id = int32(len(fs.perlinPerm) / 768)
			for k := 0; k < 256; k++ {
				push3(&fs.perlinVec, x.noise.randvec[k])
			}
			for _, perm := range [][256]int{x.noise.permX, x.noise.permY, x.noise.permZ} {
				for k := 0; k < 256; k++ {
					fs.perlinPerm = append(fs.perlinPerm, int32(perm[k]))
				}
			}
			f.perlins[x.noise] = id
		}
		even = id
	case *ImageTexture: // ImageLoader.data as the reference holds it: after its load-time LinearToGamma (rt/image_loader.go:62-70)
		if x.image == nil || x.image.Height() <= 0 {
			return 0, fmt.Errorf("flatten: ImageTexture without image data (the reference would draw its cyan debug colour)")
		}
		typ = rtxTexImage
		even = int32(len(fs.imageWidth))
		fs.imageWidth = append(fs.imageWidth, int32(x.image.imageWidth))
		fs.imageHeight = append(fs.imageHeight, int32(x.image.imageHeight))
		fs.imageOffset = append(fs.imageOffset, int64(len(fs.imageRGB)/3))
		for _, c := range x.image.data {
			push3(&fs.imageRGB, c)
		}
	default:
		return 0, fmt.Errorf("flatten: texture type %T cannot run on the device", t)
	}
	id := int32(len(fs.texType))
	fs.texType = append(fs.texType, typ)
	push3(&fs.texColor, color)
	fs.texInvScale = append(fs.texInvScale, inv)
	fs.texEven = append(fs.texEven, even)
	fs.texOdd = append(fs.texOdd, odd)
	f.texIDs[t] = id
	return id, nil
}

func (f *flattener) material(m Material) (int32, error) {
	if m == nil {
		return 0, fmt.Errorf("flatten: nil material")
	}
	if id, ok := f.matIDs[m]; ok {
		return id, nil
	}
	fs := f.fs
	typ, tex := int32(0), int32(-1)
	var albedo Color
	fuzz, ior := 0.0, 0.0
	var err error
	switch x := m.(type) {
	case *Lambertian:
		typ = rtxMatLambertian
		tex, err = f.texture(x.tex)
	case *Metal:
		typ, albedo, fuzz = rtxMatMetal, x.Albedo, x.Fuzz
	case *Dielectric:
		typ, ior = rtxMatDielectric, x.RefractionIndex
	case *DiffuseLight:
		typ = rtxMatDiffuseLight
		tex, err = f.texture(x.tex)
	case *Isotropic:
		typ = rtxMatIsotropic
		tex, err = f.texture(x.tex)
	default:
		return 0, fmt.Errorf("flatten: material type %T cannot run on the device", m)
	}
	if err != nil {
		return 0, err
	}
	id := int32(len(fs.matType))
	fs.matType = append(fs.matType, typ)
	fs.matTex = append(fs.matTex, tex)
	push3(&fs.matAlbedo, albedo)
	fs.matFuzz = append(fs.matFuzz, fuzz)
	fs.matIor = append(fs.matIor, ior)
	f.matIDs[m] = id
	return id, nil
}

// primitive returns (kind, index) of a bare primitive, adding it on first sight; ok = false when h is not one.
func (f *flattener) primitive(h Hittable) (kind, index int32, ok bool, err error) {
	fs := f.fs
	if id, seen := f.primIDs[h]; seen {
		switch h.(type) {
		case *Sphere:
			return rtxGeomSphere, id, true, nil
		case *Quad:
			return rtxGeomQuad, id, true, nil
		case *Triangle:
			return rtxGeomTriangle, id, true, nil
		case *Plane:
			return rtxGeomPlane, id, true, nil
		case *Circle:
			return rtxGeomCircle, id, true, nil
		}
	}
	var mat int32
	switch x := h.(type) {
	case *Sphere: // Center = Ray{orig: center1, dir: center2 - center1} (rt/sphere.go:17, :27); the RAW radius: the bbox uses it (rt/sphere.go:15-21)
		if mat, err = f.material(x.Mat); err != nil {
			return
		}
		kind, index = rtxGeomSphere, int32(len(fs.sphMat))
		push3(&fs.sphCenter, x.Center.Origin())
		push3(&fs.sphVelocity, x.Center.Direction())
		// Sphere stores max(0, radius); its bounding box was built from the raw value. The two only differ for negative radii, which no
		// scene function uses; a drop-in that must cover them adds an unexported rawRadius field to Sphere.
		fs.sphRadius = append(fs.sphRadius, x.Radius)
		fs.sphMat = append(fs.sphMat, mat)
	case *Quad:
		if mat, err = f.material(x.mat); err != nil {
			return
		}
		kind, index = rtxGeomQuad, int32(len(fs.quadMat))
		push3(&fs.quadQ, x.Q)
		push3(&fs.quadU, x.u)
		push3(&fs.quadV, x.v)
		fs.quadMat = append(fs.quadMat, mat)
	case *Triangle:
		if mat, err = f.material(x.mat); err != nil {
			return
		}
		kind, index = rtxGeomTriangle, int32(len(fs.triMat))
		push3(&fs.triV0, x.v0)
		push3(&fs.triV1, x.v1)
		push3(&fs.triV2, x.v2)
		fs.triMat = append(fs.triMat, mat)
		fs.triRank = append(fs.triRank, 0)
	case *Plane:
		if mat, err = f.material(x.Mat); err != nil {
			return
		}
		kind, index = rtxGeomPlane, int32(len(fs.planeMat))
		push3(&fs.planePoint, x.Point)
		push3(&fs.planeNormal, x.Normal)
		fs.planeMat = append(fs.planeMat, mat)
	case *Circle:
		if mat, err = f.material(x.mat); err != nil {
			return
		}
		kind, index = rtxGeomCircle, int32(len(fs.circleMat))
		push3(&fs.circleCenter, x.center)
		push3(&fs.circleNormal, x.normal)
		fs.circleRadius = append(fs.circleRadius, x.radius)
		fs.circleMat = append(fs.circleMat, mat)
	default:
		return 0, 0, false, nil
	}
	f.primIDs[h] = index
	return kind, index, true, nil
}

// dfs appends the objects of a reference BVH in the order BVHNode.Hit tests them (rt/bvh.go:219-239); a leaf that hangs on both
// sides of its node (rt/bvh.go:141) is taken once.
func dfs(h Hittable, out *[]Hittable) {
	switch x := h.(type) {
	case nil:
	case *BVHNode:
		if x.left != nil && x.left == x.right {
			dfs(x.left, out)
			return
		}
		dfs(x.left, out)
		dfs(x.right, out)
	case *BVHLeaf:
		*out = append(*out, x.objects...)
	default:
		*out = append(*out, h)
	}
}

func (f *flattener) listGroup(l *HittableList) (int32, error) {
	if id, ok := f.groupIDs[l]; ok {
		return id, nil
	}
	fs := f.fs
	begin := int32(len(fs.listItemKind))
	for _, o := range l.Objects {
		k, idx, ok, err := f.primitive(o)
		if err != nil {
			return 0, err
		}
		if !ok {
			return 0, fmt.Errorf("flatten: a nested HittableList may only hold primitives (Box = 6 quads, rt/primitives.go:5), found %T", o)
		}
		fs.listItemKind = append(fs.listItemKind, k)
		fs.listItemIndex = append(fs.listItemIndex, idx)
	}
	id := int32(len(fs.groupKind))
	fs.groupKind = append(fs.groupKind, rtxGeomList)
	fs.groupBegin = append(fs.groupBegin, begin)
	fs.groupCount = append(fs.groupCount, int32(len(l.Objects)))
	f.groupIDs[l] = id
	return id, nil
}

// meshGroup: the *BVHNode LoadOBJ returns (rt/obj_loader.go:109). Go's BVHNode does not keep its source slice, so the triangles
// are taken in the tree's own test order: primitive id == test-order rank, which is all exact ties need.
func (f *flattener) meshGroup(root *BVHNode) (int32, error) {
	if id, ok := f.groupIDs[root]; ok {
		return id, nil
	}
	fs := f.fs
	var order []Hittable
	dfs(root, &order)
	begin := int32(len(fs.triMat))
	for r, o := range order {
		t, ok := o.(*Triangle)
		if !ok {
			return 0, fmt.Errorf("flatten: a nested BVH must be a triangle mesh (rt/obj_loader.go:109), found %T", o)
		}
		mat, err := f.material(t.mat)
		if err != nil {
			return 0, err
		}
		push3(&fs.triV0, t.v0)
		push3(&fs.triV1, t.v1)
		push3(&fs.triV2, t.v2)
		fs.triMat = append(fs.triMat, mat)
		fs.triRank = append(fs.triRank, int32(r))
	}
	id := int32(len(fs.groupKind))
	fs.groupKind = append(fs.groupKind, rtxGeomMesh)
	fs.groupBegin = append(fs.groupBegin, begin)
	fs.groupCount = append(fs.groupCount, int32(len(order)))
	f.groupIDs[root] = id
	return id, nil
}

// entry flattens one object of world.Objects: [Volume] over [Translate][RotateY][Scale]... over a primitive / Box list / mesh.
func (f *flattener) entry(h Hittable) error {
	fs := f.fs
	volume := int32(-1)
	if v, ok := h.(*Volume); ok {
		mat, err := f.material(v.phaseFunction)
		if err != nil {
			return err
		}
		volume = int32(len(fs.volMat))
		fs.volNegInvDensity = append(fs.volNegInvDensity, v.negInvDensity)
		fs.volMat = append(fs.volMat, mat)
		h = v.boundary
	}
	xfBegin, xfCount := int32(len(fs.xfType)), int32(0)
wrappers:
	for {
		switch x := h.(type) {
		case *Translate:
			fs.xfType = append(fs.xfType, rtxXfTranslate)
			push3(&fs.xfA, x.Offset)
			push3(&fs.xfB, Vec3{})
			h = x.Obj
		case *RotateY:
			fs.xfType = append(fs.xfType, rtxXfRotateY)
			push3(&fs.xfA, Vec3{X: x.SinTheta, Y: x.CosTheta})
			push3(&fs.xfB, Vec3{})
			h = x.Obj
		case *Scale:
			fs.xfType = append(fs.xfType, rtxXfScale)
			push3(&fs.xfA, x.Factor)
			push3(&fs.xfB, x.InvFactor)
			h = x.Obj
		case *RotateX, *RotateZ:
			// their Hit back-transforms with the forward rotation while their boxes are rotated the other way (rt/transform.go:194-353):
			// what the reference renders through them depends on the BVH built over them. Refused, like the C library refuses them.
			return fmt.Errorf("flatten: %T is outside the device path", h)
		default:
			break wrappers
		}
		xfCount++
	}
	kind, index, ok, err := f.primitive(h)
	if err != nil {
		return err
	}
	if !ok {
		switch x := h.(type) {
		case *HittableList:
			kind = rtxGeomList
			index, err = f.listGroup(x)
		case *BVHNode:
			kind = rtxGeomMesh
			index, err = f.meshGroup(x)
		case *Volume:
			err = fmt.Errorf("flatten: a Volume inside a transform is outside the device path")
		default:
			err = fmt.Errorf("flatten: Hittable type %T cannot run on the device (user-defined hittables are not supported)", h)
		}
		if err != nil {
			return err
		}
	}
	fs.entryKind = append(fs.entryKind, kind)
	fs.entryIndex = append(fs.entryIndex, index)
	fs.entryXfBegin = append(fs.entryXfBegin, xfBegin)
	fs.entryXfCount = append(fs.entryXfCount, xfCount)
	fs.entryVolume = append(fs.entryVolume, volume)
	fs.entryRank = append(fs.entryRank, int32(len(fs.entryRank)))
	return nil
}

// flattenScene: world is the *HittableList of a scene function or the *BVHNode NewBVHNodeFromList built over it (main.go:77).
// For a BVH the entries are taken in the tree's test order, so entry index == test-order rank (the reference sorts the list's
// slice in place while it builds, rt/bvh.go:120-217: the insertion order is gone by then anyway).
func flattenScene(world Hittable, camera *Camera) (*flatScene, error) {
	fs := &flatScene{}
	f := &flattener{fs: fs, texIDs: map[Texture]int32{}, matIDs: map[Material]int32{}, primIDs: map[Hittable]int32{}, groupIDs: map[Hittable]int32{}, perlins: map[*Perlin]int32{}}
	var objects []Hittable
	switch w := world.(type) {
	case *HittableList:
		objects = w.Objects
	case *BVHNode:
		fs.worldIsBVH = true
		dfs(w, &objects)
	default:
		return nil, fmt.Errorf("flatten: world must be a *HittableList or the *BVHNode returned by NewBVHNodeFromList, got %T", world)
	}
	for _, o := range objects {
		if err := f.entry(o); err != nil {
			return nil, err
		}
	}
	for _, l := range camera.Lights { // Camera.Lights order matters: uniform pick by index (rt/camera.go:502-505)
		if q, ok := l.(*Quad); ok {
			_, idx, _, err := f.primitive(q)
			if err != nil {
				return nil, err
			}
			fs.lightQuad = append(fs.lightQuad, idx)
		} else {
			fs.lightQuad = append(fs.lightQuad, -1) // sampleAreaLight returns black for a non-quad light (rt/camera.go:616-619)
		}
	}
	if env := camera.Environment; env != nil && env.IsValid() {
		fs.envWidth, fs.envHeight = env.width, env.height
		fs.envRGB = make([]float64, 0, 3*env.width*env.height)
		for _, c := range env.image.data { // decoded linear pixels, row-major, y = 0 top (rt/image_loader.go:374-382)
			push3(&fs.envRGB, c)
		}
		fs.envRotation = env.rotation
		fs.envImportanceSampling = env.useImportanceSampling
	}
	return fs, nil
}

"""The Go side of the drop-in (go/rt/*.go) cannot be compiled here (no Go toolchain). What can be checked without one: every C
identifier the cgo bridge uses is declared in include/rtx_b200.h, every field of rtx_scene_desc / rtx_camera_desc is filled by
the bridge, the type switch of the flattener names every type the C++ mirror's flattener handles, and brackets balance."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "rtx_b200.h")).read()


def _struct_fields(name):
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", HDR, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\[[^\]]*\])?\s*$", part.strip())
            if m:
                fields.append(m.group(1))
    return fields


def test_cgo_identifiers_exist_in_the_header():
    files = sorted(glob.glob(os.path.join(ROOT, "go", "rt", "*.go")))
    assert [os.path.basename(f) for f in files] == ["bucket_renderer_gpu.go", "flatten.go", "gpu_bridge.go"]
    used = set()
    for f in files:
        src = open(f).read()
        for a, b in ("{}", "()", "[]"):
            assert src.count(a) == src.count(b), f"{f}: unbalanced {a}{b}"
        assert src.lstrip().startswith(("//", "package")), f
        assert re.search(r"^package rt$", src, re.M), f"{f}: must live in the reference's rt package"
        used |= set(re.findall(r"C\.(rtx_[A-Za-z0-9_]+|RTX_[A-Z0-9_]+)", src))
    assert {"rtx_create_multi", "rtx_scene_upload", "rtx_camera_set", "rtx_accum_clear", "rtx_render_pass", "rtx_resolve_rgba8",
            "rtx_get_stats", "rtx_destroy", "rtx_last_error"} <= used
    for ident in used:
        assert re.search(r"\b" + ident + r"\b", HDR), f"go/rt uses C.{ident}, which include/rtx_b200.h does not declare"


def test_bridge_fills_every_descriptor_field():
    src = open(os.path.join(ROOT, "go", "rt", "gpu_bridge.go")).read()
    for struct in ("rtx_scene_desc", "rtx_camera_desc"):
        for field in _struct_fields(struct):
            assert re.search(r"\.\s*" + field + r"\b", src), f"gpu_bridge.go never touches {struct}.{field}"


def test_flattener_type_switch_covers_the_mirror():
    go = open(os.path.join(ROOT, "go", "rt", "flatten.go")).read()
    for typ in ("*Sphere", "*Quad", "*Triangle", "*Plane", "*Circle", "*HittableList", "*BVHNode", "*Translate", "*RotateY", "*Scale", "*Volume",
                "*Lambertian", "*Metal", "*Dielectric", "*DiffuseLight", "*Isotropic", "*SolidColor", "*CheckerTexture", "*NoiseTexture", "*ImageTexture"):
        assert re.search(r"case\s+[^:]*" + re.escape(typ) + r"\b", go), f"flatten.go has no case for {typ}"
    assert "RotateX" in go and "RotateZ" in go and "error" in go    # refused, never a fallback

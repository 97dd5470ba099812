#!/usr/bin/env python3
"""Pins the scene constants to the REFERENCE SOURCE: a small interpreter for the subset of Go that
/root/reference/rt/scenes.go is written in executes every scene function of that file and records what it builds.

  python tools/extract_scene_constants.py            -> tests/golden/scenes_ref.json

Test infrastructure (it reads /root/reference, which only exists in the build container; the JSON it writes is committed
and is what tests/test_scene_constants.py compares the host mirror's flattened scenes against, field by field).

How it works: scenes.go only declares functions that call constructors of the rt package (NewQuad, NewSphere, Box,
NewTransform().Set...().Apply(), NewCameraBuilder()...Build(), world.Add(...)) with numeric literals, struct literals,
a little arithmetic, two nested loops (RandomScene) and one range loop (CornellBoxLucy). The interpreter evaluates that
code; every call of a function that scenes.go does not define itself becomes a *record* {id, fn, args}; a method call on
a record becomes a record with a receiver. Vec3 / Point3 / Color values and their Add / Sub / Len / Scale methods are
evaluated natively (rt/vec3.go). RandomDouble() draws from the SplitMix64 stream the host mirror uses for its seeded
RandomScene (seed 0x5EED; the reference itself draws from Go's auto-seeded source, so its geometry differs run to run),
in Go's evaluation order (operands left to right, struct fields in source order).
"""
import json
import math
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/rt/scenes.go"
OUT = os.path.join(ROOT, "tests", "golden", "scenes_ref.json")

# ---------------------------------------------------------------------------------------------------- tokens
TOK = re.compile(r"""
    (?P<ws>[ \t\r]+) | (?P<nl>\n) | (?P<lc>//[^\n]*) | (?P<bc>/\*.*?\*/) |
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?) | (?P<str>"(?:[^"\\]|\\.)*") |
    (?P<id>[A-Za-z_][A-Za-z_0-9]*) |
    (?P<op>:=|\+\+|--|==|!=|<=|>=|&&|\|\||[-+*/<>=!&|(){}\[\],.;:])
""", re.X | re.S)
KEYWORDS = {"package", "type", "func", "return", "for", "range", "if", "else", "var", "struct", "nil", "true", "false"}


def tokenize(src):
    out, pos, line = [], 0, 1
    def asi():
        if out and (out[-1][0] in ("num", "str", "id") or out[-1][1] in (")", "]", "}", "++", "--", "return", "nil", "true", "false")):
            out.append(("op", ";", line))
    while pos < len(src):
        m = TOK.match(src, pos)
        if not m:
            raise SyntaxError(f"line {line}: cannot tokenize {src[pos:pos + 20]!r}")
        pos = m.end()
        k = m.lastgroup
        if k == "nl":
            asi(); line += 1
        elif k == "lc":
            pass
        elif k == "bc":
            line += m.group().count("\n")
        elif k == "ws":
            pass
        else:
            t = m.group()
            out.append(("kw" if k == "id" and t in KEYWORDS else k, t, line))
    asi()
    out.append(("eof", "", line))
    return out


# ---------------------------------------------------------------------------------------------------- parser
class Parser:
    def __init__(self, toks):
        self.t, self.i, self.nolit = toks, 0, False
        self.types, self.funcs = {}, {}

    def peek(self, k=0): return self.t[self.i + k]
    def at(self, text): return self.t[self.i][1] == text and self.t[self.i][0] in ("op", "kw")
    def eat(self, text=None):
        tok = self.t[self.i]
        if text is not None and tok[1] != text:
            raise SyntaxError(f"line {tok[2]}: expected {text!r}, found {tok[1]!r}")
        self.i += 1
        return tok
    def skip_semis(self):
        while self.at(";"): self.i += 1

    # types: returned as ('named', name) | ('ptr', T) | ('slice', T) | ('struct', [field names])
    def parse_type(self):
        if self.at("*"): self.eat(); return ("ptr", self.parse_type())
        if self.at("["): self.eat("["); self.eat("]"); return ("slice", self.parse_type())
        if self.at("struct"):
            self.eat(); self.eat("{"); names = []
            self.skip_semis()
            while not self.at("}"):
                group = [self.eat()[1]]
                while self.at(","): self.eat(); group.append(self.eat()[1])
                ftype = self.parse_type()
                names += [(g, ftype) for g in group]
                self.skip_semis()
            self.eat("}")
            return ("struct", names)
        name = self.eat()[1]
        if self.at("."): self.eat(); name += "." + self.eat()[1]
        return ("named", name)

    def parse_file(self):
        self.skip_semis()
        self.eat("package"); self.eat(); self.skip_semis()
        while self.peek()[0] != "eof":
            if self.at("type"):
                self.eat(); name = self.eat()[1]; self.types[name] = self.parse_type()
            elif self.at("func"):
                self.eat(); name = self.eat()[1]; self.eat("(")
                params = []
                while not self.at(")"):
                    group = [self.eat()[1]]
                    while self.at(","): self.eat(); group.append(self.eat()[1])
                    self.parse_type(); params += group
                    if self.at(","): self.eat()
                self.eat(")")
                if self.at("("):       # result list
                    self.eat("(")
                    while not self.at(")"):
                        self.parse_type()
                        if self.at(","): self.eat()
                    self.eat(")")
                elif not self.at("{"):
                    self.parse_type()
                self.funcs[name] = (params, self.parse_block())
            else:
                tok = self.peek(); raise SyntaxError(f"line {tok[2]}: unexpected {tok[1]!r} at top level")
            self.skip_semis()

    def parse_block(self):
        self.eat("{"); stmts = []
        self.skip_semis()
        while not self.at("}"):
            stmts.append(self.parse_stmt()); self.skip_semis()
        self.eat("}")
        return stmts

    def parse_simple(self):
        lhs = [self.parse_expr()]
        while self.at(","): self.eat(); lhs.append(self.parse_expr())
        if self.at(":=") or self.at("="):
            op = self.eat()[1]
            if self.at("range"):
                self.eat(); return ("rangeassign", lhs, self.parse_expr())
            rhs = [self.parse_expr()]
            while self.at(","): self.eat(); rhs.append(self.parse_expr())
            return ("assign", op, lhs, rhs)
        if self.at("++") or self.at("--"):
            op = self.eat()[1]; return ("incdec", lhs[0], op)
        return ("expr", lhs[0])

    def parse_stmt(self):
        if self.at("return"):
            self.eat(); vals = []
            if not self.at(";"):
                vals.append(self.parse_expr())
                while self.at(","): self.eat(); vals.append(self.parse_expr())
            return ("return", vals)
        if self.at("var"):
            self.eat(); name = self.eat()[1]; self.parse_type()
            return ("var", name)
        if self.at("if"):
            self.eat(); old, self.nolit = self.nolit, True
            cond = self.parse_expr(); self.nolit = old
            then = self.parse_block(); els = None
            if self.at("else"):
                self.eat(); els = [self.parse_stmt()] if self.at("if") else self.parse_block()
            return ("if", cond, then, els)
        if self.at("for"):
            self.eat(); old, self.nolit = self.nolit, True
            init = self.parse_simple()
            if init[0] == "rangeassign":
                self.nolit = old; return ("forrange", init[1], init[2], self.parse_block())
            self.eat(";"); cond = self.parse_expr(); self.eat(";"); post = self.parse_simple()
            self.nolit = old
            return ("for", init, cond, post, self.parse_block())
        return self.parse_simple()

    PREC = [("||",), ("&&",), ("==", "!=", "<", "<=", ">", ">="), ("+", "-"), ("*", "/")]

    def parse_expr(self, level=0):
        if level == len(self.PREC): return self.parse_unary()
        left = self.parse_expr(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.PREC[level]:
            op = self.eat()[1]; left = ("bin", op, left, self.parse_expr(level + 1))
        return left

    def parse_unary(self):
        if self.at("-"): self.eat(); return ("neg", self.parse_unary())
        if self.at("+"): self.eat(); return self.parse_unary()
        if self.at("!"): self.eat(); return ("not", self.parse_unary())
        return self.parse_postfix(self.parse_primary())

    def parse_lit_body(self, typ):
        self.eat("{"); old, self.nolit = self.nolit, False
        items = []
        self.skip_semis()
        while not self.at("}"):
            if self.at("{"):                      # element with elided type (slice of structs)
                items.append((None, self.parse_lit_body(typ[1] if typ[0] == "slice" else typ)))
            else:
                e = self.parse_expr()
                if self.at(":"):
                    self.eat(); items.append((e[1], self.parse_expr()))
                else:
                    items.append((None, e))
            if self.at(","): self.eat()
            self.skip_semis()
        self.eat("}"); self.nolit = old
        return ("lit", typ, items)

    def parse_primary(self):
        tok = self.peek()
        if tok[0] == "num":
            self.eat(); return ("num", float(tok[1]) if re.search(r"[.eE]", tok[1]) else int(tok[1]))
        if tok[0] == "str":
            self.eat(); return ("str", json.loads(tok[1]))
        if self.at("("):
            self.eat(); old, self.nolit = self.nolit, False
            e = self.parse_expr(); self.nolit = old; self.eat(")"); return e
        if self.at("[") or self.at("struct"):
            typ = self.parse_type(); return self.parse_lit_body(typ)
        if tok[0] == "kw" and tok[1] in ("nil", "true", "false"):
            self.eat(); return ("const", {"nil": None, "true": True, "false": False}[tok[1]])
        if tok[0] == "id":
            self.eat(); return ("name", tok[1])
        raise SyntaxError(f"line {tok[2]}: unexpected {tok[1]!r} in expression")

    def parse_postfix(self, e):
        while True:
            if self.at("("):
                self.eat(); old, self.nolit = self.nolit, False
                args = []
                self.skip_semis()
                while not self.at(")"):
                    args.append(self.parse_expr())
                    if self.at(","): self.eat()
                    self.skip_semis()
                self.eat(")"); self.nolit = old
                e = ("call", e, args)
            elif self.at("."):
                self.eat(); e = ("sel", e, self.eat()[1])
            elif self.at("{") and not self.nolit and e[0] == "name" and e[1][:1].isupper():
                e = self.parse_lit_body(("named", e[1]))
            else:
                return e


# ---------------------------------------------------------------------------------------------------- values
class SplitMix64:      # host/rt_scenes.cpp: the generator of the seeded RandomScene
    def __init__(self, seed): self.s = seed & 0xFFFFFFFFFFFFFFFF
    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    def random_double(self): return float(self.next() >> 11) * (1.0 / 9007199254740992.0)


VEC_TYPES = ("Vec3", "Point3", "Color")


class Struct:
    def __init__(self, tname, fields): self.tname, self.f = tname, fields   # fields: ordered dict


class Record:
    count = 0
    def __init__(self, fn, args, recv=None):
        Record.count += 1
        self.id, self.fn, self.args, self.recv, self.calls = Record.count, fn, args, recv, []


class Return(Exception):
    def __init__(self, vals): self.vals = vals


class Interp:
    def __init__(self, parser, seed, package_vars=None):
        self.p, self.rng = parser, SplitMix64(seed)
        self.package_vars = package_vars or {}

    def struct_fields(self, typ):
        if typ[0] == "named":
            if typ[1] in VEC_TYPES: return [("X", None), ("Y", None), ("Z", None)]
            return self.struct_fields(self.p.types[typ[1]])
        if typ[0] == "struct": return typ[1]
        raise TypeError(f"not a struct type: {typ}")

    def zero(self, ftype):
        if ftype is None: return 0.0
        if ftype[0] == "named" and ftype[1] in ("int", "float64"): return 0 if ftype[1] == "int" else 0.0
        if ftype[0] in ("named", "struct"):
            try:
                fl = self.struct_fields(ftype)
            except (KeyError, TypeError):
                return None
            return Struct(ftype[1] if ftype[0] == "named" else "struct", {n: self.zero(t) for n, t in fl})
        return None

    def lit(self, typ, items, env):
        if typ[0] == "slice":
            return [self.eval(it[1], env) for it in items]
        fl = self.struct_fields(typ)
        vals = {n: self.zero(t) for n, t in fl}
        for k, (key, e) in enumerate(items):      # source order = Go's evaluation order
            v = self.lit(e[1], e[2], env) if e[0] == "lit" else self.eval(e, env)
            vals[key if key is not None else fl[k][0]] = v
        return Struct(typ[1] if typ[0] == "named" else "struct", vals)

    def call_func(self, name, args):
        params, body = self.p.funcs[name]
        env = dict(zip(params, args))
        try:
            self.exec_block(body, env)
        except Return as r:
            return r.vals[0] if len(r.vals) == 1 else tuple(r.vals)
        return None

    def vec_method(self, v, m, args):
        x, y, z = v.f["X"], v.f["Y"], v.f["Z"]
        mk = lambda a, b, c: Struct(v.tname, {"X": a, "Y": b, "Z": c})
        if m == "Add": o = args[0]; return mk(x + o.f["X"], y + o.f["Y"], z + o.f["Z"])
        if m == "Sub": o = args[0]; return mk(x - o.f["X"], y - o.f["Y"], z - o.f["Z"])
        if m == "Scale": return mk(args[0] * x, args[0] * y, args[0] * z)
        if m == "Neg": return mk(-x, -y, -z)
        if m == "Len2": return x * x + y * y + z * z
        if m == "Len": return math.sqrt(x * x + y * y + z * z)   # rt/vec3.go: math.Sqrt(Len2())
        raise NotImplementedError(f"Vec3.{m}")

    def eval(self, e, env):
        k = e[0]
        if k == "num" or k == "str" or k == "const": return e[1]
        if k == "name":
            if e[1] in env: return env[e[1]]
            if e[1] in self.package_vars: return self.package_vars[e[1]]
            raise NameError(e[1])
        if k == "neg": return -self.eval(e[1], env)
        if k == "not": return not self.eval(e[1], env)
        if k == "bin":
            a = self.eval(e[2], env)
            if e[1] == "&&": return a and self.eval(e[3], env)
            if e[1] == "||": return a or self.eval(e[3], env)
            b = self.eval(e[3], env)
            op = e[1]
            if op == "+": return a + b
            if op == "-": return a - b
            if op == "*": return a * b
            if op == "/": return a // b if isinstance(a, int) and isinstance(b, int) else a / b
            if op == "<": return a < b
            if op == "<=": return a <= b
            if op == ">": return a > b
            if op == ">=": return a >= b
            if op == "==": return a is b if (a is None or b is None) else a == b
            if op == "!=": return a is not b if (a is None or b is None) else a != b
        if k == "lit": return self.lit(e[1], e[2], env)
        if k == "sel":
            base = self.eval(e[1], env)
            if isinstance(base, Struct): return base.f[e[2]]
            raise TypeError(f"selector .{e[2]} on {type(base).__name__}")
        if k == "call":
            fn = e[1]
            if fn[0] == "sel":                                   # method call (receiver evaluated first)
                recv = self.eval(fn[1], env)
                args = [self.eval(a, env) for a in e[2]]
                if isinstance(recv, Struct) and recv.tname in VEC_TYPES: return self.vec_method(recv, fn[2], args)
                if isinstance(recv, Record):
                    r = Record("." + fn[2], args, recv)
                    recv.calls.append(r)
                    return r
                raise TypeError(f"method {fn[2]} on {type(recv).__name__}")
            name = fn[1]
            args = [self.eval(a, env) for a in e[2]]
            if name == "float64": return float(args[0])
            if name == "int": return int(args[0])
            if name == "RandomDouble": return self.rng.random_double()                          # rt/utils.go:18
            if name == "RandomDoubleRange": return args[0] + (args[1] - args[0]) * self.rng.random_double()   # rt/utils.go:22-24
            if name == "panic": raise RuntimeError(f"panic({args[0]!r})")
            if name in self.p.funcs: return self.call_func(name, args)
            r = Record(name, args)
            if name in ("LoadOBJ", "LoadOBJWithTransform"): return (r, None)     # (mesh, err)
            return r
        raise NotImplementedError(k)

    def assign(self, target, v, env):
        if target[0] == "name":
            if target[1] != "_": env[target[1]] = v
        elif target[0] == "sel": self.eval(target[1], env).f[target[2]] = v
        else: raise NotImplementedError(target[0])

    def exec_block(self, stmts, env):
        for s in stmts: self.exec(s, env)

    def exec(self, s, env):
        k = s[0]
        if k == "expr": self.eval(s[1], env)
        elif k == "var": env[s[1]] = None
        elif k == "assign":
            vals = [self.eval(r, env) for r in s[3]]
            if len(s[2]) > 1 and len(vals) == 1: vals = list(vals[0])
            for t, v in zip(s[2], vals): self.assign(t, v, env)
        elif k == "incdec": self.assign(s[1], self.eval(s[1], env) + (1 if s[2] == "++" else -1), env)
        elif k == "return": raise Return([self.eval(v, env) for v in s[1]])
        elif k == "if":
            if self.eval(s[1], env): self.exec_block(s[2], env)
            elif s[3] is not None: self.exec_block(s[3], env)
        elif k == "for":
            self.exec(s[1], env)
            while self.eval(s[2], env):
                self.exec_block(s[4], env); self.exec(s[3], env)
        elif k == "forrange":
            seq = self.eval(s[2], env)
            for idx, item in enumerate(seq):
                if len(s[1]) > 0: self.assign(s[1][0], idx, env)
                if len(s[1]) > 1: self.assign(s[1][1], item, env)
                self.exec_block(s[3], env)
        else: raise NotImplementedError(k)


# ---------------------------------------------------------------------------------------------------- output
def to_json(v, seen):
    if isinstance(v, Struct):
        if v.tname in VEC_TYPES: return [v.f["X"], v.f["Y"], v.f["Z"]]
        return {n: to_json(x, seen) for n, x in v.f.items()}
    if isinstance(v, Record):
        if v.id in seen: return {"ref": v.id}
        seen.add(v.id)
        d = {"id": v.id, "fn": v.fn, "args": [to_json(a, seen) for a in v.args]}
        if v.recv is not None: d["recv"] = to_json(v.recv, seen)
        return d
    if isinstance(v, (list, tuple)): return [to_json(x, seen) for x in v]
    return v


SCENES = {   # host-mirror scene name (rt_scenes.cpp: scene_named) -> scene function of rt/scenes.go
    "random": "RandomScene", "checkered": "CheckeredSpheresScene", "simple": "SimpleScene", "earth": "EarthScene", "perlin": "PerlinSpheresScene",
    "quads": "QuadsScene", "primitives": "PrimitivesScene", "hdri-test": "HDRITestScene", "cornell": "CornellBoxScene",
    "glossy-metal": "GlossyMetalTest", "cornell-glossy": "CornellBoxGlossy", "cornell-lucy": "CornellBoxLucy", "cornell-smoke": "CornellSmoke",
}


def package_colors(path):
    """Package-level `Name = Color{X: a, Y: b, Z: c}` variables of another file of the package (rt/camera.go:528-536: the
    Background* presets scenes.go refers to)."""
    out = {}
    for m in re.finditer(r"^\s*(\w+)\s*=\s*Color\{X:\s*([-\d.]+),\s*Y:\s*([-\d.]+),\s*Z:\s*([-\d.]+)\}", open(path).read(), re.M):
        out[m.group(1)] = Struct("Color", {"X": float(m.group(2)), "Y": float(m.group(3)), "Z": float(m.group(4))})
    return out


def run(ref=REF, seed=0x5EED):
    p = Parser(tokenize(open(ref).read()))
    p.parse_file()
    pkg = package_colors(os.path.join(os.path.dirname(ref), "camera.go"))
    out = {"_source": "rt/scenes.go of the reference, executed by tools/extract_scene_constants.py", "_seed": seed, "scenes": {}}
    for key, fn in SCENES.items():
        Record.count = 0
        it = Interp(p, seed, pkg)
        world, camera = it.call_func(fn, [])
        seen = set()
        adds = [to_json(c.args[0], seen) for c in world.calls if c.fn == ".Add"]
        out["scenes"][key] = {"function": fn, "world": adds, "camera": to_json(camera, seen), "random_draws": None}
    return out


if __name__ == "__main__":
    res = run()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    for k, v in res["scenes"].items():
        print(f"{k:15s} {v['function']:24s} {len(v['world'])} world.Add calls")
    print("wrote", OUT)
    sys.exit(0)

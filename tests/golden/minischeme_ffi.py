"""minischeme_ffi — what tests/golden/minischeme.py needs on top of itself to EXECUTE THE GUILE GLUE of this
repository (guile/griddy/cuda.scm, guile/griddy/simulate-cuda.scm) without Guile and without a GPU.

TEST INFRASTRUCTURE (tests/test_guile_glue.py).  It adds to an Interp:
  - (rnrs bytevectors): bytevectors over Python bytearrays, native little-endian accessors;
  - (system foreign): `dynamic-link`, `dynamic-func`, `pointer->procedure`, `bytevector->pointer`,
    `pointer->bytevector`, `make-pointer`, `pointer->string`, `%null-pointer`, `sizeof`, the type names.  A
    foreign procedure does not call into a library: it hands (symbol name, arguments) to a Python callback, which
    plays libgriddy_cuda.so — it records what the glue passes and fills the output buffers;
  - `do`, `values` / `call-with-values`, vectors, `getenv`, `nan?`, and inert stand-ins for the few Chickadee
    painters simulate-cuda.scm's `draw` uses.
What this proves: the Scheme glue's own logic (index conventions, array layouts, argument order, read-back) does
what INTEGRATION.md says.  What it cannot: Guile's FFI marshalling itself.
"""
from __future__ import annotations

import math
import os
import struct

import minischeme as ms
from minischeme import NIL, S, UNSPEC, Pair, Prim, Sym, plist, slist


class ByteVector:
    __slots__ = ("data",)

    def __init__(self, n, fill=0):
        self.data = bytearray([fill & 0xFF]) * n if n else bytearray()


class Pointer:
    """A foreign pointer: to a bytevector (obj), to a C string (obj is str), null (obj None), or a bare address."""
    __slots__ = ("obj", "address")

    def __init__(self, obj=None, address=0):
        self.obj, self.address = obj, address


class Values:
    __slots__ = ("items",)

    def __init__(self, items):
        self.items = list(items)


_FMT = {"u8": ("<B", 1), "s32": ("<i", 4), "s64": ("<q", 8), "f64": ("<d", 8), "f32": ("<f", 4)}


def install(I, foreign_call):
    """foreign_call(name: str, args: list) -> return value of the C function."""
    B = I.builtins

    def prim(name, fn):
        B[S(name)] = Prim(name, fn)

    # ---- (rnrs bytevectors) -----------------------------------------------------------------------
    prim("make-bytevector", lambda n, fill=0: ByteVector(int(n), int(fill)))
    prim("bytevector-length", lambda bv: len(bv.data))
    prim("native-endianness", lambda: S("little"))

    def setter(kind):
        fmt, _ = _FMT[kind]
        return lambda bv, at, v: (struct.pack_into(fmt, bv.data, int(at), float(v) if kind[0] == "f" else int(v)), UNSPEC)[1]

    def getter(kind):
        fmt, _ = _FMT[kind]
        return lambda bv, at: struct.unpack_from(fmt, bv.data, int(at))[0]
    for scheme, kind in (("u8", "u8"), ("s32-native", "s32"), ("s64-native", "s64"), ("ieee-double-native", "f64"),
                         ("ieee-single-native", "f32")):
        prim(f"bytevector-{scheme}-set!", setter(kind))
        prim(f"bytevector-{scheme}-ref", getter(kind))
    prim("bytevector-uint-ref", lambda bv, at, endian, size: int.from_bytes(bv.data[int(at): int(at) + int(size)], "little"))

    # ---- (system foreign) -------------------------------------------------------------------------
    for t in ("int", "int64", "double", "void", "size_t", "uint8", "int32"):
        B[S(t)] = S(t)
    B[S("%null-pointer")] = Pointer(None)
    prim("sizeof", lambda t: {"*": 8, "int": 4, "int64": 8, "double": 8, "size_t": 8, "uint8": 1, "int32": 4}[t.name])
    prim("dynamic-link", lambda name=None: ("dynamic-library", name))
    prim("dynamic-func", lambda name, lib: name)
    prim("bytevector->pointer", lambda bv, offset=0: Pointer(bv))
    prim("make-pointer", lambda address: Pointer(None, int(address)))
    prim("null-pointer?", lambda p: p.obj is None and p.address == 0)

    def pointer_to_string(p):
        if not isinstance(p.obj, str):
            raise ms.SchemeError("pointer->string on something that is not a C string")
        return p.obj
    prim("pointer->string", pointer_to_string)

    def pointer_to_bytevector(p, n):
        if not isinstance(p.obj, ByteVector) or len(p.obj.data) < int(n):
            raise ms.SchemeError("pointer->bytevector: not that much memory behind the pointer")
        return p.obj
    prim("pointer->bytevector", pointer_to_bytevector)

    def pointer_to_procedure(ret, name, arg_types):
        n_args = len(plist(arg_types))

        def call(*args):
            if len(args) != n_args:
                raise ms.SchemeError(f"{name}: bound with {n_args} arguments, called with {len(args)}")
            return foreign_call(name, list(args))
        return Prim(name, call)
    prim("pointer->procedure", pointer_to_procedure)

    # ---- language bits the glue uses that the reference does not ---------------------------------------
    def sf_do(I_, x, env):
        specs = plist(x.cdr.car)
        test = plist(x.cdr.cdr.car)
        body = x.cdr.cdr.cdr
        loop = ms.Env(env)
        for sp in specs:
            loop.vars[sp.car] = I_.eval(sp.cdr.car, env)
        while True:
            if I_.eval(test[0], loop) is not False:
                r = UNSPEC
                for e in test[1:]:
                    r = I_.eval(e, loop)
                return r
            inner = ms.Env(loop)
            for form in plist(body):          # (eval_body would hand back the last form as a tail call)
                I_.eval(form, inner)
            new = {}
            for sp in specs:
                if sp.cdr.cdr is not NIL:
                    new[sp.car] = I_.eval(sp.cdr.cdr.car, loop)
            nxt = ms.Env(env)
            nxt.vars.update(loop.vars)
            nxt.vars.update(new)
            loop = nxt
    ms.SPECIAL["do"] = sf_do

    def sf_with_style(I_, x, env):      # chickadee (graphics path): styling is irrelevant here
        return I_.eval_body(x.cdr.cdr, env)
    ms.SPECIAL["with-style"] = sf_with_style

    prim("values", lambda *a: a[0] if len(a) == 1 else Values(a))

    def call_with_values(producer, consumer):
        v = I.apply(producer, [])
        return I.apply(consumer, v.items if isinstance(v, Values) else [v])
    prim("call-with-values", call_with_values)

    prim("list->vector", lambda l: list(plist(l)))
    prim("vector-ref", lambda v, i: v[int(i)])
    prim("vector-length", lambda v: len(v))
    prim("getenv", lambda name: os.environ.get(name, False))
    prim("nan?", lambda v: isinstance(v, float) and math.isnan(v))
    prim("hashq-set!", lambda t, k, v: I.apply(I.builtins[S("hash-table-set!")], [t, k, v]))
    prim("hashq-ref", lambda t, k, d=False: I.apply(I.builtins[S("hash-table-ref/default")], [t, k, d]))

    # ---- what guile/tools/headless-reference.scm uses ------------------------------------------------
    import re as _re
    I.stdout, I.stderr, I.argv = [], [], ["guile"]
    ERR = ("port", "stderr")
    prim("current-error-port", lambda: ERR)
    prim("current-output-port", lambda: ("port", "stdout"))

    def scheme_repr(v, display=True):
        if isinstance(v, float):
            return repr(v)
        if isinstance(v, str):
            return v if display else '"' + v + '"'
        if v is True:
            return "#t"
        if v is False:
            return "#f"
        return repr(v)

    def fmt(dest, template, *args):
        args = list(args)

        def sub(m):
            d = m.group(0)
            if d == "~%":
                return "\n"
            if d == "~~":
                return "~"
            v = args.pop(0)
            if d in ("~a", "~A"):
                return scheme_repr(v)
            if d in ("~s", "~S"):
                return scheme_repr(v, display=False)
            return f"{float(v):.{int(m.group(1))}f}"          # ~,Nf
        out = _re.sub(r"~,(\d+)f|~[aAsS%~]", sub, template)
        if dest is False:
            return out
        (I.stderr if dest == ERR else I.stdout).append(out)
        return UNSPEC
    prim("format", fmt)
    prim("command-line", lambda: slist(I.argv))
    prim("string->number", lambda t: int(t) if _re.fullmatch(r"[+-]?\d+", t) else float(t))
    prim("cadar", lambda l: l.car.cdr.car)
    prim("get-internal-real-time", lambda: 0)
    B[S("internal-time-units-per-second")] = 1000000000
    prim("version", lambda: "minischeme (not Guile)")

    class SchemeExit(Exception):
        pass
    I.SchemeExit = SchemeExit

    def scheme_exit(code=0):
        raise SchemeExit(code)
    prim("exit", scheme_exit)
    prim("resolve-module", lambda name: I.module(name))

    def module_ref(module, sym):
        return I.lookup(sym, module.env)

    def module_set(module, sym, value):
        module.env.vars[sym] = value
        I.version += 1
        return UNSPEC
    prim("module-ref", module_ref)
    prim("module-set!", module_set)
    prim("load", lambda path: (I.load_file(path, I.user), UNSPEC)[1])

    # painters used by simulate-cuda.scm's `draw`
    B[S("*actor/color*")] = S("actor-colour")
    prim("circle", lambda centre, radius: ("circle", centre.x, centre.y, radius))
    prim("fill", lambda shape: shape)
    prim("superimpose", lambda *shapes: ("superimpose", list(shapes)))
    prim("make-canvas", lambda painter: ("canvas", painter))

    def draw_canvas(canvas):
        I.drawn.append(canvas)
        return UNSPEC
    I.drawn = []
    prim("draw-canvas", draw_canvas)
    return I

"""
ORACLE (test infrastructure, not product code): run the UNMODIFIED reference in this container.

The reference (/root/reference) is Python 2 and JIT-compiles its inner loops with
``weave.inline``; neither Python 2, ``weave``, ``linop`` nor ``krypy`` exist in this image.  This
loader makes the reference's own source text executable under Python 3 WITHOUT copying it:

* the four modules on the hot path are read from /root/reference at run time, passed through a
  purely lexical py2->py3 pass (``print`` statements, ``xrange``, tab expansion) and exec'd;
* ``weave.inline`` is served by a small shim that compiles the reference's *actual C++ loop
  bodies* with g++ -O3 (the blitz ``a(i)`` indexing is mapped onto raw pointers with macros) and
  calls them through ctypes, picking the variables out of the caller's frame as weave does;
* ``linop`` is served by oracle/linop_min.py (operator algebra only, no arithmetic);
* ``krypy`` is a stub: the two functions that call it cannot be run (parity unpinned there).

It is used only by tests/golden/make_golden.py to generate the golden vectors committed under
tests/golden/ -- /root/reference does not exist on the GPU box, so nothing at test/bench time
imports this module.
"""
import ctypes
import hashlib
import os
import re
import subprocess
import sys
import tempfile
import types

import numpy as np

REF_ROOT = os.environ.get("COSMOMAP2_REFERENCE", "/root/reference")
_CACHE = os.path.join(tempfile.gettempdir(), "cm2_refrun_cache")

# ---------------------------------------------------------------------------------------------
# mini-weave
# ---------------------------------------------------------------------------------------------
_CT = {
    np.dtype(np.float64): ("double", ctypes.c_double),
    np.dtype(np.float32): ("float", ctypes.c_float),
    np.dtype(np.int64): ("long", ctypes.c_long),
    np.dtype(np.int32): ("int", ctypes.c_int),
    np.dtype(np.bool_): ("unsigned char", ctypes.c_ubyte),
    np.dtype(np.uint8): ("unsigned char", ctypes.c_ubyte),
}
_loaded = {}


def _inline(code, arg_names=(), local_dict=None, global_dict=None, support_code="", **kw):
    frame = sys._getframe(1)
    loc = frame.f_locals if local_dict is None else local_dict
    glob = frame.f_globals if global_dict is None else global_dict
    params, macros, cargs, keep = [], [], [], []
    sig = []
    for name in arg_names:
        val = loc[name] if name in loc else glob[name]
        if isinstance(val, np.ndarray):
            if val.ndim != 1:
                raise NotImplementedError("mini-weave: only 1-d arrays (%s)" % name)
            cname, ctyp = _CT[val.dtype]
            stride = val.strides[0] // val.itemsize if val.size else 1
            params.append("%s* %s__p, long %s__s" % (cname, name, name))
            macros.append("#define %s(i) %s__p[(long)(i)*%s__s]" % (name, name, name))
            cargs += [ctypes.c_void_p(val.ctypes.data), ctypes.c_long(stride)]
            keep.append(val)
            sig.append("A:" + cname)
        elif isinstance(val, (bool, int, np.integer)):
            params.append("long %s" % name)
            cargs.append(ctypes.c_long(int(val)))
            sig.append("L")
        elif isinstance(val, (float, np.floating)):
            params.append("double %s" % name)
            cargs.append(ctypes.c_double(float(val)))
            sig.append("D")
        else:
            raise NotImplementedError("mini-weave: unsupported %s=%r" % (name, type(val)))
    # support_code only carries #include lines in the reference; omp.h is harmless
    src = "#include <math.h>\n#include <stdio.h>\n#include <stdlib.h>\n%s\n%s\n" \
          "extern \"C\" double cm2_weave_fn(%s){\n double return_val=0;\n%s\n return return_val;\n}\n" \
          % (support_code, "\n".join(macros), ", ".join(params), code)
    key = hashlib.sha1((src + "|".join(sig)).encode()).hexdigest()
    fn = _loaded.get(key)
    if fn is None:
        os.makedirs(_CACHE, exist_ok=True)
        so = os.path.join(_CACHE, key + ".so")
        if not os.path.exists(so):
            cpp = os.path.join(_CACHE, key + ".cpp")
            with open(cpp, "w") as f:
                f.write(src)
            subprocess.check_call(["g++", "-O3", "-fopenmp", "-fPIC", "-shared", "-w", "-o", so, cpp])
        lib = ctypes.CDLL(so)
        fn = lib.cm2_weave_fn
        fn.restype = ctypes.c_double
        _loaded[key] = fn
    return fn(*cargs)


def _make_weave():
    weave = types.ModuleType("weave")
    weave.inline = _inline
    weave.converters = types.SimpleNamespace(blitz="blitz")
    return weave


# ---------------------------------------------------------------------------------------------
# lexical py2 -> py3
# ---------------------------------------------------------------------------------------------
_PRINT = re.compile(r"^(\s*)print\s+(?!\()(.*)$")
_PRINT_PAREN_PCT = re.compile(r"^(\s*)print\s*(\(.*\)\s*%.*)$")


def _py3(src):
    out = []
    for line in src.expandtabs(8).split("\n"):
        m = _PRINT.match(line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        else:
            m = _PRINT_PAREN_PCT.match(line)
            if m:
                line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = re.sub(r"\bxrange\b", "range", line)
        out.append(line)
    return "\n".join(out)


def _exec_module(name, relpath, namespace_extra=None, drop_imports=()):
    path = os.path.join(REF_ROOT, relpath)
    with open(path) as f:
        src = f.read()
    lines = []
    for line in _py3(src).split("\n"):
        s = line.strip()
        if any(s.startswith(d) for d in drop_imports):
            line = line[:len(line) - len(line.lstrip())] + "pass"
        lines.append(line)
    mod = types.ModuleType(name)
    mod.__file__ = path
    if namespace_extra:
        mod.__dict__.update(namespace_extra)
    sys.modules[name] = mod
    exec(compile("\n".join(lines), path, "exec"), mod.__dict__)
    return mod


_ref = None


def load_reference(quiet=True):
    """Return a namespace holding the reference's own classes/functions for the hot path."""
    global _ref
    if _ref is not None:
        return _ref
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree %s not present" % REF_ROOT)
    from . import linop_min

    saved = {k: sys.modules.get(k) for k in
             ("weave", "linop", "linop.linop", "krypy", "blkop", "utilities",
              "utilities_functions", "linear_algebra_funcs", "process_ces")}
    sys.modules["weave"] = _make_weave()
    linop_pkg = types.ModuleType("linop")
    linop_pkg.__dict__.update({k: getattr(linop_min, k) for k in dir(linop_min) if not k.startswith("__")})
    linop_pkg.linop = linop_min
    sys.modules["linop"] = linop_pkg
    sys.modules["linop.linop"] = linop_min
    krypy = types.ModuleType("krypy")
    sys.modules["krypy"] = krypy
    try:
        uf = _exec_module("utilities_functions", "utilities/utilities_functions.py")
        la = _exec_module("linear_algebra_funcs", "utilities/linear_algebra_funcs.py")
        pc = _exec_module("process_ces", "utilities/process_ces.py")
        util = types.ModuleType("utilities")
        for m in (uf, la, pc):
            util.__dict__.update({k: v for k, v in m.__dict__.items() if not k.startswith("__")})
        util.__all__ = [k for k in util.__dict__ if not k.startswith("_")]
        sys.modules["utilities"] = util
        blk = _exec_module("blkop", "interfaces/blkop.py")
        lo = _exec_module("cm2ref_linearoperators", "interfaces/linearoperators.py")
        dl = _exec_module("cm2ref_deflationlib", "interfaces/deflationlib.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    ns = types.SimpleNamespace()
    for m in (uf, la, pc, blk, lo, dl):
        for k, v in m.__dict__.items():
            if not k.startswith("__"):
                setattr(ns, k, v)
    ns.lp = linop_min
    if quiet:
        _silence(pc)
        _silence(lo)
        _silence(dl)
    _ref = ns
    return ns


def _silence(mod):
    mod.__dict__["print"] = lambda *a, **k: None

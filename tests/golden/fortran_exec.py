"""Mechanical Fortran -> Python translator for the numerical routines of the reference's physics/zm_conv.F90.

Purpose: the reference cannot be compiled in this image (no Fortran compiler), so the CPU
oracle is a hand restatement.  To pin that restatement to the reference's OWN SOURCE TEXT, this script reads a
routine from /root/reference/physics/zm_conv.F90, translates it statement by statement into Python (same
statements, same order, same operator precedence; IEEE doubles, glibc log/exp like a gfortran build), and executes
it.  `make_reference_fixtures.py` runs the translated routines on seeded inputs and commits inputs + outputs under
tests/golden/; tests then compare the oracle with those fixtures.  Nothing here is hand-written physics: the only
hand-written numerics are the externals that are NOT in the reference tree (qsat_water etc., see externals below).

Supported subset (what the ZM routines use): free-form source, `&` continuations, `!` comments, declarations
(real(r8)/integer/logical with intent/dimension/parameter/pointer, with or without `::`), assignments (scalars,
array elements, whole arrays / full-slice sections), block and one-line IF, DO with optional step and construct
names, EXIT/CYCLE, CALL (scalar intent(out)/(inout) dummies are returned and re-assigned at the call site), RETURN,
function results, the intrinsics abs min max log log10 exp sqrt sign merge nint int real mod.  WRITE/FORMAT/USE/
IMPLICIT are skipped, `call endrun` raises.  Arrays are 1-based (`FArr`).
"""
from __future__ import annotations

import math
import re

import numpy as np


class FortranStop(RuntimeError):
    pass


class FStruct:
    """Derived-type variable: components are attributes created by ALLOCATE / by the caller."""


class FArr:
    """1-based Fortran array (column-major semantics are irrelevant here: only element access is used)."""

    def __init__(self, shape, dtype=float, data=None, lb=None):
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.lb = tuple(lb) if lb is not None else (1,) * len(self.shape)
        self.a = np.zeros(self.shape, dtype=dtype) if data is None else data
        if data is None and dtype is float:
            self.a[...] = FArr.UNDEFINED  # uninitialised reals must not silently look like zeros

    UNDEFINED = np.nan                    # value of never-assigned local reals (Fortran: undefined)

    def _ix(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        out = []
        for d, i in enumerate(idx):
            if isinstance(i, slice):
                lo = None if i.start is None else i.start - self.lb[d]
                hi = None if i.stop is None else i.stop - self.lb[d] + 1      # Fortran upper bound inclusive
                out.append(slice(lo, hi))
            else:
                j = int(i) - self.lb[d]
                if j < 0 or j >= self.shape[d]:
                    raise IndexError(f"Fortran index {i} out of bounds for dimension {d + 1} ({self.shape})")
                out.append(j)
        return tuple(out)

    def __getitem__(self, idx):
        v = self.a[self._ix(idx)]
        if isinstance(v, np.ndarray):
            return v
        # reals stay numpy float64 (IEEE semantics for x/0 instead of Python's ZeroDivisionError); integers become int
        return int(v) if self.a.dtype.kind == "i" else v

    def __setitem__(self, idx, val):
        self.a[self._ix(idx)] = val

    def setall(self, val):
        self.a[...] = val.a if isinstance(val, FArr) else val

    # whole-array operands of elementwise expressions (`qfac = 1.0_r8/qfac`, geopotential.F90:263): the result is
    # the numpy array of the elementwise IEEE operation, assigned back through setall()
    @staticmethod
    def _v(x):
        return x.a if isinstance(x, FArr) else x

    def __add__(self, o): return self.a + FArr._v(o)
    def __radd__(self, o): return FArr._v(o) + self.a
    def __sub__(self, o): return self.a - FArr._v(o)
    def __rsub__(self, o): return FArr._v(o) - self.a
    def __mul__(self, o): return self.a * FArr._v(o)
    def __rmul__(self, o): return FArr._v(o) * self.a

    def __truediv__(self, o):
        with np.errstate(divide="ignore", invalid="ignore"):
            return self.a / FArr._v(o)

    def __rtruediv__(self, o):
        with np.errstate(divide="ignore", invalid="ignore"):
            return FArr._v(o) / self.a


def _ipow(x, n):
    """x**n for a literal integer n the way GCC's powi expansion does it (square and multiply, left to right)."""
    if n == 0:
        return 1.0
    if n == 1:
        return x
    if n == 2:
        return x * x
    if n == 3:
        return (x * x) * x
    if n == 4:
        t = x * x
        return t * t
    r = _ipow(x, n // 2)
    r = r * r
    return r * x if n % 2 else r


def _sign(a, b):
    return math.copysign(abs(a), b)


def _merge(a, b, c):
    return a if c else b


def _nint(x):
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def _frange(a, b, s=1):
    a, b, s = int(a), int(b), int(s)
    return range(a, b + (1 if s > 0 else -1), s)


def _log(x):
    if x == 0.0:
        return -math.inf
    if x < 0.0 or x != x:
        return math.nan
    return math.log(x)


def _log10(x):
    if x == 0.0:
        return -math.inf
    if x < 0.0 or x != x:
        return math.nan
    return math.log10(x)


def _fdiv(a, b):
    """IEEE division (Python raises on x/0.0)."""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


INTRINSICS = {
    "abs": "_abs", "min": "_min", "max": "_max", "log": "_log", "log10": "_log10", "exp": "math.exp",
    "sqrt": "math.sqrt", "sign": "_sign", "merge": "_merge", "nint": "_nint", "int": "int", "real": "_real",
    "mod": "math.fmod", "dble": "float", "minval": "_minval", "maxval": "_maxval", "present": "_present",
    "sum": "_sum", "size": "_size", "allocated": "_present", "associated": "_present", "_arr": "_arr",
}


def _arr(*a):
    return list(a)


def _elementwise(args):
    return any(isinstance(a, np.ndarray) for a in args)


def _min(*a):
    if _elementwise(a):
        r = a[0]
        for b in a[1:]:
            r = np.minimum(r, b)
        return r
    return min(a)


def _max(*a):
    if _elementwise(a):
        r = a[0]
        for b in a[1:]:
            r = np.maximum(r, b)
        return r
    return max(a)


def _abs(x):
    return np.abs(x) if isinstance(x, np.ndarray) else abs(x)


def _minval(a):
    return a.min().item() if hasattr(a, "min") else a


def _maxval(a):
    return a.max().item() if hasattr(a, "max") else a


def _present(x):
    return x is not None


def _sum(a):
    # Fortran SUM of a section: sequential left-to-right accumulation is what gfortran -O2 (no fast-math) emits
    tot = 0.0
    for v in np.asarray(a).ravel(order="F"):
        tot = tot + v
    return tot


def _size(a, dim=None):
    sh = a.shape if isinstance(a, FArr) else np.asarray(a).shape
    return int(np.prod(sh)) if dim is None else sh[dim - 1]


def _real(x, kind=None):
    return float(x)


# ------------------------------------------------------------------------------------------ source handling
def read_source(path):
    with open(path) as f:
        return f.read().split("\n")


def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def logical_lines(lines):
    """Comments stripped, continuation lines joined, lower-cased outside strings; yields (lineno, text)."""
    buf, start = "", None
    for n, raw in enumerate(lines, 1):
        s = _strip_comment(raw).strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if start is None:
            start = n
        if s.endswith("&"):
            buf += s[:-1] + " "
            continue
        buf += s
        yield start, _lower(buf)
        buf, start = "", None


def _components(s):
    """Derived-type component references a%b become plain identifiers a__pct__b (outside strings)."""
    out, q, i = [], None, 0
    parts = re.split(r"('[^']*'|\"[^\"]*\")", s)
    for k, p in enumerate(parts):
        if k % 2 == 0:
            p = re.sub(r"\s*%\s*", "__pct__", p)
        out.append(p)
    return "".join(out)


def _lower(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


ROUTINE_START = re.compile(r"^(?:(real\(r8\)|integer|logical)\s+)?(function|subroutine)\s+(\w+)\s*\(([^)]*)\)")
ROUTINE_END = re.compile(r"^end\s*(function|subroutine)(\s+\w+)?$")


def extract_routine(lines, name):
    """Returns the logical lines [(lineno, text)] of routine `name` (header ... end) and its first/last line."""
    name = name.lower()
    out, inside = [], False
    for n, text in logical_lines(lines):
        m = ROUTINE_START.match(text)
        if not inside and m and m.group(3) == name:
            inside = True
        if inside:
            out.append((n, text))
            if ROUTINE_END.match(text):
                return out
    raise KeyError(name)


# ------------------------------------------------------------------------------------------ declarations
DECL = re.compile(r"^(real\s*\(\s*r8\s*\)|real|integer|logical|character\s*\([^)]*\)|type\s*\([^)]*\))\s*(.*)$")


def _split_top(s, sep=","):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


class Var:
    def __init__(self, name, ftype):
        self.name, self.ftype = name, ftype
        self.dims = None          # list of dimension expression strings, or None for scalars
        self.intent = None
        self.parameter = False
        self.init = None
        self.pointer = False
        self.optional = False


def parse_decl(text):
    """Returns list of Var for a declaration line, or None if the line is not a declaration."""
    m = DECL.match(text)
    if not m:
        return None
    ftype, rest = m.group(1).replace(" ", ""), m.group(2).strip()
    if rest.startswith("function"):
        return None
    attrs, ents = "", rest
    if "::" in rest:
        attrs, ents = rest.split("::", 1)
    elif rest.startswith(","):
        raise SyntaxError("attribute list without '::' : " + text)
    attr_list = [a.strip() for a in _split_top(attrs.strip().lstrip(","))] if attrs.strip() else []
    dims_attr, intent, param, pointer, optional = None, None, False, False, False
    for a in attr_list:
        if a.startswith("dimension"):
            dims_attr = _split_top(a[a.index("(") + 1:a.rindex(")")])
        elif a.startswith("intent"):
            intent = a[a.index("(") + 1:a.rindex(")")].replace(" ", "")
        elif a == "parameter":
            param = True
        elif a in ("pointer", "allocatable", "target", "save", "optional"):
            pointer = pointer or a in ("pointer", "allocatable")
            optional = optional or a == "optional"
    out = []
    for e in _split_top(ents):
        init = None
        if "=" in e and "=>" not in e:
            e, init = e.split("=", 1)
            e, init = e.strip(), init.strip()
        mm = re.match(r"^(\w+)\s*(?:\((.*)\))?$", e.strip())
        if not mm:
            raise SyntaxError("cannot parse entity '%s' in: %s" % (e, text))
        v = Var(mm.group(1), ftype)
        v.dims = _split_top(mm.group(2)) if mm.group(2) else (list(dims_attr) if dims_attr else None)
        v.intent, v.parameter, v.init, v.pointer = intent, param, init, pointer
        v.optional = optional
        out.append(v)
    return out


# ------------------------------------------------------------------------------------------ expressions
TOKEN = re.compile(r"""\s*(?:
      (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_r8|_\w+)?)
    | (?P<dotop>\.(?:and|or|not|true|false|eq|ne|gt|ge|lt|le|eqv|neqv)\.)
    | (?P<name>[a-z_]\w*)
    | (?P<str>'[^']*'|"[^"]*")
    | (?P<op>\*\*|==|/=|<=|>=|=>|[-+*/(),<>=:%])
    )""", re.X)

DOTOPS = {".and.": " and ", ".or.": " or ", ".not.": " not ", ".true.": " True ", ".false.": " False ",
          ".eq.": " == ", ".ne.": " != ", ".gt.": " > ", ".ge.": " >= ", ".lt.": " < ", ".le.": " <= ",
          ".eqv.": " == ", ".neqv.": " != "}


def tokenize(s):
    s = s.replace("(/", " _arr( ").replace("/)", " ) ")      # array constructor
    pos, out = 0, []
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            if s[pos:].strip() == "":
                break
            raise SyntaxError("cannot tokenize: %r at %r" % (s, s[pos:pos + 20]))
        pos = m.end()
        kind = m.lastgroup
        out.append((kind, m.group(kind)))
    return out


class Translator:
    def __init__(self, arrays, functions, result_name=None, float_div=True):
        self.arrays = arrays            # set of array names visible in this routine
        self.functions = functions      # set of callable user routine names
        self.result_name = result_name
        self.float_div = float_div

    def expr(self, s):
        toks = tokenize(s)
        out, _ = self._seq(toks, 0, stop=None)
        return out.strip()

    def _seq(self, toks, i, stop):
        """Translate tokens until a top-level token in `stop`; returns (python, next index)."""
        parts = []
        while i < len(toks):
            kind, t = toks[i]
            if stop and kind == "op" and t in stop:
                break
            if kind == "num":
                parts.append(self._number(t))
                i += 1
            elif kind == "dotop":
                parts.append(DOTOPS[t])
                i += 1
            elif kind == "str":
                parts.append(t)
                i += 1
            elif kind == "name" and i + 1 < len(toks) and toks[i + 1] == ("op", "%"):
                # derived-type component reference a%b[(idx)][%c[(idx)]...] -> attribute access on an FStruct;
                # indexed components are arrays
                ref = t
                i += 1
                while i < len(toks) and toks[i] == ("op", "%"):
                    ref += "." + toks[i + 1][1]
                    i += 2
                    if i < len(toks) and toks[i] == ("op", "("):
                        args, i = self._args(toks, i + 1)
                        ref += "[%s]" % ", ".join(a if a != "" else ":" for a in args)
                parts.append(ref)
            elif kind == "name":
                if i + 1 < len(toks) and toks[i + 1] == ("op", "("):
                    args, i = self._args(toks, i + 2)
                    if t in self.arrays:
                        parts.append("%s[%s]" % (t, ", ".join(a if a != "" else ":" for a in args)))
                    elif t in INTRINSICS:
                        parts.append("%s(%s)" % (INTRINSICS[t], ", ".join(args)))
                    elif t in self.functions:
                        parts.append("%s(%s)" % (t, ", ".join(args)))
                    else:
                        raise NameError("unknown array/function '%s'" % t)
                else:
                    parts.append("_ret" if t == self.result_name else t)
                    i += 1
            elif kind == "op" and t == "**" and i + 1 < len(toks) and toks[i + 1][0] == "num" \
                    and re.match(r"^\d+$", toks[i + 1][1]) and parts:
                # integer power: compilers expand x**n by repeated multiplication (x**2 = x*x, x**3 = (x*x)*x,
                # x**4 = (x*x)*(x*x)), not by a call to pow()
                base = parts.pop()
                parts.append("_ipow(%s, %s)" % (base.strip(), toks[i + 1][1]))
                i += 2
            elif kind == "op":
                if t == "/=":
                    parts.append(" != ")
                elif t == "(":
                    inner, i2 = self._seq(toks, i + 1, stop=(")",))
                    parts.append("(" + inner + ")")
                    i = i2
                elif t == ":":
                    parts.append(":")
                elif t == "/":
                    parts.append(" /FDIV/ ")
                else:
                    parts.append(" " + t + " " if t not in ("(", ")") else t)
                i += 1
            else:
                raise SyntaxError(str(toks[i]))
        return "".join(parts), i

    @staticmethod
    def _number(t):
        """Fortran literal -> Python.  Integers stay integers.  Reals: `_r8` kind or a `d` exponent is double; a real
        literal with NO kind suffix is default (single) precision -- its value is the float32 rounding of the decimal
        text, promoted to double when it meets an r8 operand (zm_conv.F90:3680 has such a literal, 0.85)."""
        m = re.match(r"^(.*?)(?:_(\w+))?$", t)
        body, kind = m.group(1), m.group(2)
        is_real = ("." in body) or ("e" in body) or ("d" in body)
        if not is_real:
            return body
        if "d" in body:
            return body.replace("d", "e")
        if kind is not None:
            return body
        return repr(float(np.float32(float(body))))

    def _args(self, toks, i):
        """toks[i] is the first token after '('.  Returns (list of python arg strings, index after ')')."""
        args = []
        while True:
            a, i = self._seq(toks, i, stop=(",", ")"))
            a = a.strip()
            args.append(self._slice(a))
            if i >= len(toks):
                raise SyntaxError("unbalanced parentheses")
            if toks[i] == ("op", ")"):
                return args, i + 1
            i += 1

    @staticmethod
    def _slice(a):
        # Fortran section bounds lo:hi are inclusive; FArr handles that, we only need valid python slice syntax
        return a


def fix_div(py):
    """Replace the /FDIV/ markers: true division; integer/integer division does not occur in the routines this
    translator is used for (checked by the caller through `assert_no_integer_division`)."""
    return py.replace("/FDIV/", "/")


# ------------------------------------------------------------------------------------------ routine translation
SKIP_STMT = re.compile(r"^(use\s|implicit\s|write\s*\(|print\s|\d+\s+format|format\s*\(|save\b|external\b|intrinsic\b)")
SKIP_CALLS = {"outfld", "t_startf", "t_stopf"}


class Routine:
    def __init__(self, name, kind, args, vars_, body, result_type):
        self.name, self.kind, self.args, self.vars, self.body, self.result_type = name, kind, args, vars_, body, result_type

    def out_scalars(self):
        """Scalar dummies the caller must re-assign (intent out/inout or unspecified-intent scalars that are assigned)."""
        outs = []
        for a in self.args:
            v = self.vars.get(a)
            if v is None or v.dims is not None or v.pointer or v.ftype.startswith("type"):
                continue
            if v.intent in ("out", "inout") or (v.intent is None and self._assigned(a)):
                outs.append(a)
        return outs

    def _assigned(self, name):
        pat = re.compile(r"^(?:\w+\s*:\s*)?(?:if\s*\(.*\)\s*)?%s\s*=[^=]" % re.escape(name))
        return any(pat.match(t) for _, t in self.body)


def parse_routine(llines):
    n0, head = llines[0]
    m = ROUTINE_START.match(head)
    kind, name = m.group(2), m.group(3)
    args = [a.strip() for a in m.group(4).split(",") if a.strip()]
    vars_, body, in_decl = {}, [], True
    for n, text in llines[1:-1]:
        if SKIP_STMT.match(text):
            continue
        if in_decl:
            d = parse_decl(text)
            if d is not None:
                for v in d:
                    vars_[v.name] = v
                continue
            in_decl = False
        body.append((n, text))
    return Routine(name, kind, args, vars_, body, m.group(1))


class Module:
    """Holds translated routines and the exec namespace."""

    def __init__(self, src_path, namespace, lenient=False):
        self.lenient = lenient
        self.lines = read_source(src_path)
        self.ns = dict(namespace)
        self.ns.update(dict(math=math, np=np, FArr=FArr, FStruct=FStruct, _ipow=_ipow, _arr=_arr, r8=8, _min=_min, _max=_max,
                            _abs=_abs, _sign=_sign, _merge=_merge, _nint=_nint, _frange=_frange,
                            _log=_log, _log10=_log10, _real=_real, FortranStop=FortranStop, _minval=_minval,
                            _maxval=_maxval, _present=_present, _sum=_sum, _size=_size))
        self.routines = {}
        self.py = {}
        self._pending_goto = []
        self._goto_depth = 0

    def module_parameters(self, first_line, last_line):
        """Evaluate the module-level declarations with initialisers (parameters and initialised variables) found in
        source lines [first_line, last_line] into the namespace."""
        tr = Translator(set(), set())
        for n, text in logical_lines(self.lines[first_line - 1:last_line]):
            d = parse_decl(text)
            if not d:
                continue
            for v in d:
                if v.init is not None:
                    self.ns[v.name] = eval(fix_div(tr.expr(v.init)), self.ns)

    def run_lines(self, first_line, last_line, arrays=()):
        """Translate and execute the executable statements of source lines [first_line, last_line] in the module
        namespace (used for isolated glue statements of routines that are not translated as a whole)."""
        arr = set(arrays) | {k for k, v in self.ns.items() if isinstance(v, FArr)}
        tr = Translator(arr, {k for k, v in self.ns.items() if callable(v) and not k.startswith("_")})
        out, ind, loops = [], 0, []
        for n, text in logical_lines(self.lines[first_line - 1:last_line]):
            ind = self._stmt(text, tr, emit_ind=lambda s, i=None: out.append("    " * (ind if i is None else i) + s),
                             ind=ind, loops=loops, ret="pass", routine=None)
        src = "\n".join(out)
        exec(compile(src, "<fortran:lines %d-%d>" % (first_line, last_line), "exec"), self.ns)
        return src

    def load(self, *names):
        for name in names:
            self.routines[name.lower()] = parse_routine(extract_routine(self.lines, name))
        for name in names:
            self._translate(self.routines[name.lower()])

    # ---- code generation
    def _translate(self, r):
        arrays = {v.name for v in r.vars.values() if v.dims is not None}
        arrays |= {k for k, v in self.ns.items() if isinstance(v, FArr)}
        funcs = set(self.routines) | {k for k, v in self.ns.items() if callable(v) and not k.startswith("_")}
        tr = Translator(arrays, funcs, result_name=r.name if r.kind == "function" else None)
        sig = [a + ("=None" if (a in r.vars and r.vars[a].optional) else "") for a in r.args]
        out = ["def %s(%s):" % (r.name, ", ".join(sig))]
        ind = 1

        def emit(s):
            out.append("    " * ind + s)

        # locals: arrays allocated, parameters evaluated
        for v in r.vars.values():
            if v.name in r.args:
                continue
            if v.ftype.startswith("type"):
                emit("%s = FStruct()" % v.name)
                continue
            if v.ftype.startswith("character"):
                continue
            if v.dims is not None and (v.pointer or any(d.strip() == ":" for d in v.dims)):
                continue                                   # allocated later (or never used)
            if v.dims is not None:
                dims = ", ".join(fix_div(tr.expr(d)) for d in v.dims)
                emit("%s = FArr((%s,), dtype=%s)" % (v.name, dims, "float" if v.ftype.startswith("real") else "int"))
            elif v.init is not None:
                emit("%s = %s" % (v.name, fix_div(tr.expr(v.init))))
            elif v.name != r.name:
                # Fortran leaves locals undefined; NaN / a sentinel makes any use-before-definition visible
                emit("%s = %s" % (v.name, "math.nan" if v.ftype.startswith("real") else
                                  ("False" if v.ftype == "logical" else "-2147483647")))
        # module variables assigned by this routine (zm_convi sets the module's private data)
        local_names = set(r.vars) | set(r.args)
        assigned = set()
        for _, text in r.body:
            mm = re.match(r"^(?:if\s*\(.*\)\s*)?([a-z_]\w*)\s*=(?!=)", text)
            if mm and mm.group(1) not in local_names and mm.group(1) != r.name:
                assigned.add(mm.group(1))
        if assigned:
            out.insert(1, "    global " + ", ".join(sorted(assigned)))
        if r.kind == "function":
            emit("_ret = None")
        outs = r.out_scalars()
        ret = "return " + ("_ret" if r.kind == "function" else ("(" + ", ".join(outs) + ("," if len(outs) == 1 else "") + ")" if outs else "None"))
        loops = []                       # stack of construct names (or None)
        for n, text in r.body:
            try:
                ind = self._stmt(text, tr, emit_ind=lambda s, i=None: out.append("    " * (ind if i is None else i) + s),
                                 ind=ind, loops=loops, ret=ret, routine=r)
            except Exception as e:       # noqa: BLE001
                if not self.lenient or re.match(r"^(end\s*(if|do)|else|if\s*\(.*\)\s*then|do\s)", text):
                    raise type(e)("%s (line %d of the reference: %s)" % (e, n, text)) from e
                # lenient mode: a statement outside the supported subset is kept as a run-time stop, never skipped
                out.append("    " * ind + "raise FortranStop(%r)" % ("untranslated statement (line %d): %s" % (n, text)))
        out.append("    " + ret)
        src = "\n".join(out)
        self.py[r.name] = src
        exec(compile(src, "<fortran:%s>" % r.name, "exec"), self.ns)

    def _stmt(self, text, tr, emit_ind, ind, loops, ret, routine):
        def emit(s, i=None):
            emit_ind(s, ind if i is None else i)

        t = text.strip()
        if SKIP_STMT.match(t):          # e.g. the statement of a one-line IF is a WRITE
            emit("pass")
            return ind
        # construct name prefix  "name: do ..."
        cname = None
        m = re.match(r"^(\w+)\s*:\s*(do\b.*)$", t)
        if m:
            cname, t = m.group(1), m.group(2)
        if re.match(r"^end\s*if$", t):
            return ind - 1
        m = re.match(r"^end\s*do(\s+\w+)?$", t)
        if m:
            if self._pending_goto and len(loops) == self._goto_depth:
                self._expect_label = self._pending_goto[-1]      # next statement must be `<label> continue`
            loops.pop()
            return ind - 1
        if getattr(self, "_expect_label", None):
            lab, self._expect_label = self._expect_label, None
            if not re.match(r"^%s\s+continue$" % lab, t):
                raise NotImplementedError("GOTO %s does not target the statement after its loop" % lab)
        if t == "else":
            emit("else:", ind - 1)
            emit("pass")
            return ind
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", t)
        if m:
            emit("elif %s:" % fix_div(tr.expr(m.group(1))), ind - 1)
            emit("pass")
            return ind
        m = re.match(r"^if\s*\((.*)\)\s*then$", t)
        if m:
            emit("if %s:" % fix_div(tr.expr(m.group(1))))
            emit("pass", ind + 1)
            return ind + 1
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", t)
        if m:
            parts = _split_top(m.group(2))
            args = ", ".join(fix_div(tr.expr(p)) for p in parts)
            emit("for %s in _frange(%s):" % (m.group(1), args))
            emit("pass", ind + 1)
            loops.append(cname)
            return ind + 1
        if t.startswith("if"):
            # one-line IF: find the matching parenthesis of the condition
            k = t.index("(")
            depth, j = 0, k
            while True:
                if t[j] == "(":
                    depth += 1
                elif t[j] == ")":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            cond, rest = t[k + 1:j], t[j + 1:].strip()
            emit("if %s:" % fix_div(tr.expr(cond)))
            self._stmt(rest, tr, emit_ind, ind + 1, loops, ret, routine)
            return ind
        m = re.match(r"^exit(\s+(\w+))?$", t)
        if m:
            if m.group(2) and (not loops or loops[-1] != m.group(2)):
                raise NotImplementedError("EXIT of a non-innermost loop")
            emit("break")
            return ind
        if t == "cycle":
            emit("continue")
            return ind
        if t == "return":
            emit(ret)
            return ind
        if t == "continue":
            emit("pass")
            return ind
        m = re.match(r"^(\d+)\s+continue$", t)
        if m:
            if self._pending_goto and self._pending_goto[-1] == m.group(1):
                self._pending_goto.pop()
            emit("pass")
            return ind
        m = re.match(r"^go\s*to\s+(\d+)$", t)
        if m:
            # only the pattern `goto L` ... `end do` / `L continue` (= EXIT of the innermost loop) is accepted; the
            # label check happens when `L continue` is met right after that loop
            if not loops:
                raise NotImplementedError("GOTO outside a loop")
            self._pending_goto.append(m.group(1))
            self._goto_depth = len(loops)
            emit("break")
            return ind
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", t)
        if m:
            name, argtxt = m.group(1), m.group(2) or ""
            if name == "endrun":
                emit("raise FortranStop(%s)" % (fix_div(tr.expr(argtxt)) or "'endrun'"))
                return ind
            if name in SKIP_CALLS and name not in self.ns:       # history output etc.; a shim in the namespace captures it
                emit("pass")
                return ind
            args = [fix_div(tr.expr(a)) for a in _split_top(argtxt)]
            callee = self.routines.get(name)
            if callee is not None:
                outs = callee.out_scalars()
                pos = [callee.args.index(o) for o in outs]
                pure_out = [callee.args.index(o) for o in outs if callee.vars[o].intent == "out"]
            else:
                pos = list(self.ns.get("_OUTS", {}).get(name, []))
                pure_out = pos
            targets = [args[i] for i in pos]
            args = [("None" if i in pure_out else a) for i, a in enumerate(args)]     # intent(out): no value goes in
            call = "%s(%s)" % (name, ", ".join(args))
            if targets:
                emit("%s = %s" % (", ".join(targets) + ("," if len(targets) == 1 else ""), call))
            else:
                emit(call)
            return ind
        m = re.match(r"^allocate\s*\((.*)\)$", t)
        if m:
            for item in _split_top(m.group(1)):
                mm = re.match(r"^([\w%\s]+?)\s*\((.*)\)$", item.strip())
                if not mm:
                    continue                               # stat= etc.
                dims = ", ".join(fix_div(tr.expr(d)) for d in _split_top(mm.group(2)))
                emit("%s = FArr((%s,), dtype=float)" % (re.sub(r"\s*%\s*", ".", mm.group(1)), dims))
            return ind
        if re.match(r"^(deallocate|nullify)\s*\(", t):
            emit("pass")
            return ind
        # assignment (pointer assignment => is not used by the routines in scope)
        m = re.match(r"^([a-z_]\w*)\s*(%|\(|=(?!=))", t)
        if m and "=>" not in t:
            lhs, rhs = self._split_assign(t)
            rhs_py = fix_div(tr.expr(rhs))
            if "%" in lhs:
                mm = re.match(r"^(.*?)\s*(?:\(([^()]*)\))?$", lhs)
                base, idx = mm.group(1), mm.group(2)
                if idx is None or all(p.strip() == ":" for p in _split_top(idx)):
                    emit("%s.setall(%s)" % (fix_div(tr.expr(base)), rhs_py))
                else:
                    emit("%s = %s" % (fix_div(tr.expr(lhs)), rhs_py))
                return ind
            mm = re.match(r"^([a-z_]\w*)\s*(?:\((.*)\))?$", lhs)
            name, idx = mm.group(1), mm.group(2)
            if name in tr.arrays:
                if idx is None or all(p.strip() == ":" for p in _split_top(idx)):
                    emit("%s.setall(%s)" % (name, rhs_py))
                else:
                    emit("%s = %s" % (fix_div(tr.expr(lhs)), rhs_py))
            else:
                emit("%s = %s" % ("_ret" if name == tr.result_name else name, rhs_py))
            return ind
        raise NotImplementedError("statement not supported: " + t)

    @staticmethod
    def _split_assign(t):
        depth = 0
        for k, ch in enumerate(t):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and t[k + 1] != "=" and t[k - 1] not in "<>/=":
                return t[:k].strip(), t[k + 1:].strip()
        raise SyntaxError("no assignment in: " + t)

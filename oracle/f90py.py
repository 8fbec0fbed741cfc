"""Run the reference's Fortran sources WHERE THEY LIE, without a Fortran compiler (TEST INFRASTRUCTURE).

The image has no gfortran, so `oracle/_ref` cannot be built from /root/reference/f90/*.f90.  This module is the
substitute: a mechanical translator of the Fortran 90 subset those files use (free-form source, declarations with
explicit bounds, do / if / cycle, whole-array and array-section expressions with strides, allocatable locals, the
intrinsics below, FFTW's plan / execute calls, OpenMP regions incl. the thread-number-driven chunk deposits with
their barriers) into Python/numpy source, executed statement by statement on numpy arrays.  Nothing here knows what
any subroutine computes: signs, index offsets, operation order, the reference's quirks all come from the Fortran
text.  tools/gen_golden_f90.py uses it to record golden vectors for every hot-path subroutine
(tests/golden/f90_*.npz); tests/test_f90_golden.py checks the C++ oracle and, on the GPU box, the CUDA library
against them.  It reads /root/reference and therefore only runs in the build container.

What is NOT the reference's own arithmetic: FFTW3 is a third-party library absent from the reference tree; its calls
are served by numpy.fft with FFTW's published convention (unnormalised, forward e^{-i..}, backward e^{+i..}).
`-ffast-math` re-association (the reference's build flag) is not emulated: expressions are evaluated in source order
in IEEE double precision.
"""
import os
import re

import numpy as np

# ------------------------------------------------------------------------------------------------ runtime


class S:
    """array section lo:hi:step (Fortran: inclusive bounds); None = the array's own bound"""

    __slots__ = ("lo", "hi", "st")

    def __init__(self, lo=None, hi=None, st=None):
        self.lo, self.hi, self.st = lo, hi, st


class FArr:
    """numpy array with Fortran lower bounds; sections come back as plain numpy views"""

    __slots__ = ("d", "lb")

    def __init__(self, d, lb):
        self.d, self.lb = d, tuple(int(x) for x in lb)
        assert d.ndim == len(self.lb)

    def _key(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        assert len(idx) == self.d.ndim, ("rank mismatch", len(idx), self.d.ndim)
        key = []
        for ax, (i, lb) in enumerate(zip(idx, self.lb)):
            n = self.d.shape[ax]
            if isinstance(i, S):
                st = 1 if i.st is None else int(i.st)
                lo = (lb if st > 0 else lb + n - 1) if i.lo is None else int(i.lo)
                hi = (lb + n - 1 if st > 0 else lb) if i.hi is None else int(i.hi)
                a, b = lo - lb, hi - lb
                if st > 0:
                    if a < 0 or b >= n:
                        raise IndexError("section %d:%d outside bounds %d:%d (axis %d)" % (lo, hi, lb, lb + n - 1, ax))
                    key.append(slice(a, b + 1, st))
                else:
                    if b < 0 or a >= n:
                        raise IndexError("section %d:%d:%d outside bounds (axis %d)" % (lo, hi, st, ax))
                    key.append(slice(a, b - 1 if b - 1 >= 0 else None, st))
            else:
                j = int(i) - lb
                if j < 0 or j >= n:
                    raise IndexError("index %d outside bounds %d:%d (axis %d)" % (int(i), lb, lb + n - 1, ax))
                key.append(j)
        return tuple(key)

    def __getitem__(self, idx):
        return self.d[self._key(idx)]

    def __setitem__(self, idx, val):
        self.d[self._key(idx)] = val

    def copy(self):
        return FArr(self.d.copy(), self.lb)


def _div(a, b):
    """Fortran '/': integer operands truncate toward zero"""
    ia = isinstance(a, (int, np.integer)) or (isinstance(a, np.ndarray) and a.dtype.kind == "i")
    ib = isinstance(b, (int, np.integer)) or (isinstance(b, np.ndarray) and b.dtype.kind == "i")
    if ia and ib:
        q = np.trunc(np.true_divide(a, b))
        return q.astype(np.int64) if isinstance(q, np.ndarray) else int(q)
    return a / b


def _pow(a, b):
    if isinstance(b, (int, np.integer)) and not isinstance(a, np.ndarray) and isinstance(a, (int, np.integer)):
        return int(a) ** int(b)
    return a ** b


def _floor(x):
    r = np.floor(x)
    return r.astype(np.int64) if isinstance(r, np.ndarray) else int(r)


def _int(x, kind=None):
    r = np.trunc(np.real(x))
    return r.astype(np.int64) if isinstance(r, np.ndarray) else int(r)


def _nint(x):
    r = np.sign(x) * np.floor(np.abs(x) + 0.5)
    return r.astype(np.int64) if isinstance(r, np.ndarray) else int(r)


def _cmplx(a, b=0.0, kind=None):
    return a + 1j * b if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else complex(a, b)


def _r4(x):
    """a literal without a d-exponent is REAL(4)"""
    return float(np.float32(x))


def _sum(x, dim=None):
    if dim is not None:  # SUM(array, DIM=d): the other axes are kept; element order along d as in memory
        return np.add.reduce(np.asarray(x), axis=int(dim) - 1)
    # sequential left-to-right like a Fortran loop (numpy's pairwise sum only differs for > 8 elements)
    x = np.asarray(x).ravel(order="F")
    acc = x[0] * 0
    for v in x:
        acc = acc + v
    return acc


def _sign(a, b):
    return np.abs(a) if b >= 0 else -np.abs(a)


def _fftw_plan(n, a, b, sign, flags):
    return (int(sign), int(n))


def _fftw_exec(plan, a, b):
    sign, n = plan
    src = a.d if isinstance(a, FArr) else a
    dst = b.d if isinstance(b, FArr) else b
    dst[...] = np.fft.fft(src) if sign < 0 else np.fft.ifft(src) * n


def _run_team(region, nthreads):
    """an OpenMP team whose members run to their next barrier in turn (the regions that need this index their work
    by omp_get_thread_num and only touch shared data another member also touches after a barrier)"""
    gens = [region(t) for t in range(int(nthreads))]
    live = list(gens)
    while live:
        nxt = []
        for g in live:
            try:
                next(g)
                nxt.append(g)
            except StopIteration:
                pass
        live = nxt


RUNTIME = dict(
    np=np, S=S, FArr=FArr, _div=_div, _pow=_pow, _floor=_floor, _int=_int, _nint=_nint, _cmplx=_cmplx, _r4=_r4, _sum=_sum,
    _sign=_sign, _fftw_plan=_fftw_plan, _fftw_exec=_fftw_exec, _run_team=_run_team,
    fftw_forward=-1, fftw_backward=1, fftw_estimate=64, fftw_destroy_input=1, fftw_measure=0,
)

INTRINSICS = {
    "sqrt": "np.sqrt", "dsqrt": "np.sqrt", "abs": "np.abs", "dabs": "np.abs", "cdabs": "np.abs", "exp": "np.exp",
    "dexp": "np.exp", "cdexp": "np.exp", "cos": "np.cos", "dcos": "np.cos", "sin": "np.sin", "dsin": "np.sin",
    "tan": "np.tan", "atan": "np.arctan", "datan": "np.arctan", "atan2": "np.arctan2", "datan2": "np.arctan2",
    "cosh": "np.cosh", "dcosh": "np.cosh", "sinh": "np.sinh", "dsinh": "np.sinh", "log": "np.log", "dlog": "np.log",
    "floor": "_floor", "int": "_int", "nint": "_nint", "dble": "np.real", "real": "np.real", "aimag": "np.imag",
    "dimag": "np.imag", "conjg": "np.conj", "dconjg": "np.conj", "cmplx": "_cmplx", "dcmplx": "_cmplx", "sum": "_sum",
    "mod": "np.fmod", "min": "min", "max": "max", "maxval": "np.max", "minval": "np.min", "sign": "_sign", "fftw_plan_dft_1d": "_fftw_plan",
}

# ------------------------------------------------------------------------------------------------ lexer / parser
TOK = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[de][+-]?\d+)?(?:_\w+)?)
  | (?P<dotop>\.(?:and|or|not|eq|ne|lt|le|gt|ge|true|false|eqv|neqv)\.)
  | (?P<name>[a-z_]\w*)
  | (?P<op>\*\*|==|/=|<=|>=|//|[-+*/(),:<>=%])
)""", re.X)


def tokenize(src):
    pos, out = 0, []
    src = src.strip()
    while pos < len(src):
        m = TOK.match(src, pos)
        if not m or m.end() == pos:
            raise SyntaxError("cannot tokenize %r at %d" % (src, pos))
        pos = m.end()
        kind = m.lastgroup
        text = m.group(kind)
        if kind == "num" and out and out[-1] == ("dotop_pending", None):
            pass
        out.append((kind, text))
    # "1.and." style collisions do not occur in the reference; guard against silent mis-lexing of e.g. "2.eq.x"
    return out


class Parser:
    def __init__(self, toks, unit):
        self.t, self.i, self.u = toks, 0, unit

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def eat(self, text=None):
        kind, tx = self.peek()
        if text is not None and tx != text:
            raise SyntaxError("expected %r, got %r in %r" % (text, tx, self.t))
        self.i += 1
        return kind, tx

    def done(self):
        return self.i >= len(self.t)

    # precedence: .or. < .and. < .not. < comparison < +- < */ < unary - < **
    def expr(self):
        a = self.and_()
        while self.peek()[1] in (".or.",):
            self.eat()
            a = "(%s or %s)" % (a, self.and_())
        return a

    def and_(self):
        a = self.not_()
        while self.peek()[1] == ".and.":
            self.eat()
            a = "(%s and %s)" % (a, self.not_())
        return a

    def not_(self):
        if self.peek()[1] == ".not.":
            self.eat()
            return "(not %s)" % self.not_()
        return self.cmp()

    CMP = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", ".eq.": "==", ".ne.": "!=", ".lt.": "<",
           ".le.": "<=", ".gt.": ">", ".ge.": ">="}

    def cmp(self):
        a = self.add()
        if self.peek()[1] in self.CMP:
            op = self.CMP[self.eat()[1]]
            a = "(%s %s %s)" % (a, op, self.add())
        return a

    def add(self):
        if self.peek()[1] in ("+", "-"):
            sg = self.eat()[1]
            a = self.mul()
            a = "(-%s)" % a if sg == "-" else a
        else:
            a = self.mul()
        while self.peek()[1] in ("+", "-"):
            op = self.eat()[1]
            a = "(%s %s %s)" % (a, op, self.mul())
        return a

    def mul(self):
        a = self.unary()
        while self.peek()[1] in ("*", "/"):
            op = self.eat()[1]
            b = self.unary()
            a = "(%s * %s)" % (a, b) if op == "*" else "_div(%s, %s)" % (a, b)
        return a

    def unary(self):
        if self.peek()[1] == "-":
            self.eat()
            return "(-%s)" % self.unary()
        if self.peek()[1] == "+":
            self.eat()
            return self.unary()
        return self.power()

    def power(self):
        a = self.primary()
        if self.peek()[1] == "**":
            self.eat()
            b = self.unary()  # right associative, binds tighter than unary minus on the left
            return "_pow(%s, %s)" % (a, b)
        return a

    def number(self, text):
        t = re.sub(r"_\w+$", "", text)
        if re.fullmatch(r"\d+", t):
            return t
        if "d" in t:
            return "%r" % float(t.replace("d", "e"))
        v = float(t)
        if float(np.float32(v)) != v:
            self.u.warnings.append("single-precision literal %s is not exact" % text)
        return "_r4(%r)" % v

    def primary(self):
        kind, tx = self.eat()
        if kind == "num":
            return self.number(tx)
        if kind == "dotop":
            return {".true.": "True", ".false.": "False"}[tx]
        if tx == "(":
            a = self.expr()
            if self.peek()[1] == ",":  # complex literal (re, im)
                self.eat(",")
                b = self.expr()
                self.eat(")")
                return "complex(%s, %s)" % (a, b)
            self.eat(")")
            return "(%s)" % a
        if kind == "name":
            if self.peek()[1] == "(":
                self.eat("(")
                args = self.arglist()
                self.eat(")")
                if tx in self.u.arrays:
                    return "%s[%s]" % (self.u.py(tx), ", ".join(args) + ("," if len(args) == 1 else ""))
                if tx == "omp_get_thread_num":
                    return "_tid"
                if tx in INTRINSICS:
                    if any(a.startswith("S(") for a in args):
                        raise SyntaxError("section passed to intrinsic %s" % tx)
                    return "%s(%s)" % (INTRINSICS[tx], ", ".join(args))
                raise SyntaxError("unknown function or undeclared array %r" % tx)
            if tx in self.u.arrays:
                return "%s.d" % self.u.py(tx)
            if tx == "omp_get_thread_num":
                return "_tid"
            return self.u.py(tx)
        raise SyntaxError("unexpected token %r in %r" % (tx, self.t))

    def arglist(self):
        args = []
        if self.peek()[1] in (")", None):
            return args
        while True:
            if self.peek()[0] == "name" and self.peek(1)[1] == "=" :  # keyword argument (SUM(x, DIM=1))
                kw = self.eat()[1]
                self.eat("=")
                args.append("%s=%s" % (kw, self.expr()))
            else:
                args.append(self.section_or_expr())
            if self.peek()[1] == ",":
                self.eat(",")
                continue
            return args

    def section_or_expr(self):
        parts, cur, is_sec = [], None, False
        if self.peek()[1] != ":":
            cur = self.expr()
        parts.append(cur)
        while self.peek()[1] == ":":
            self.eat(":")
            is_sec = True
            cur = None
            if self.peek()[1] not in (":", ",", ")", None):
                cur = self.expr()
            parts.append(cur)
        if not is_sec:
            return parts[0]
        # keyword form keeps "S(" at the start so that callers can recognise sections
        return "S(%s)" % ", ".join("None" if p is None else p for p in parts)


# ------------------------------------------------------------------------------------------------ translator
PYKW = {"in", "lambda", "is", "not", "and", "or", "if", "else", "for", "while", "def", "class", "from", "import", "pass",
        "del", "global", "with", "as", "try", "except", "raise", "return", "yield", "print", "len", "np", "S", "type"}


class Unit:
    def __init__(self, name, args):
        self.name, self.args = name, args
        self.arrays = {}     # name -> (dtype, [(lo, hi)] | None for allocatable)
        self.scalars = {}    # name -> dtype char: i / r / r4 / c / p (pointer-like)
        self.inits = []      # (name, python rhs)
        self.warnings = []
        self.body = []

    @staticmethod
    def py(name):
        return name + "_" if name in PYKW else name


def split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return [x.strip() for x in out]


def logical_lines(text):
    """comments stripped ('!$omp' kept), continuation lines joined, lower-cased; yields (line_no, text)"""
    buf, start, omp = "", None, False
    for no, raw in enumerate(text.splitlines(), 1):
        line = raw.rstrip()
        st = line.strip()
        low = st.lower()
        if low.startswith("!$omp"):
            body = low[5:].strip()
            if buf and omp:  # continued directive ("!$omp private(a,&" / "!$omp   b)")
                buf += " " + body.lstrip("&").strip()
            else:
                if buf:
                    yield start, buf
                buf, start, omp = "!$omp " + body, no, True
            if buf.endswith("&"):
                buf = buf[:-1].rstrip()
                continue
            yield start, buf
            buf, omp = "", False
            continue
        if "!" in line:  # no string literals with '!' in the reference's f90 files (checked by the translator's user)
            line = line[:line.index("!")].rstrip()
        st = line.strip().lower()
        if not st:
            continue
        if buf:
            buf += " " + st.lstrip("&").strip()
        else:
            buf, start = st, no
        if buf.endswith("&"):
            buf = buf[:-1].rstrip()
            continue
        for piece in split_top(buf, ";"):
            if piece:
                yield start, piece
        buf = ""
    if buf:
        yield start, buf


DECL = re.compile(r"^(integer|real|complex|double precision|type|logical)\b")


class Translator:
    def __init__(self, path):
        self.path = path
        self.units = {}
        lines = list(logical_lines(open(path).read()))
        cur = None
        for no, ln in lines:
            m = re.match(r"^subroutine\s+(\w+)\s*\((.*)\)\s*$", ln)
            if m:
                cur = Unit(m.group(1), [a.strip() for a in m.group(2).split(",") if a.strip()])
                self.units[cur.name] = cur
                continue
            if re.match(r"^end(\s+subroutine.*)?$", ln):
                cur = None
                continue
            if cur is not None:
                cur.body.append((no, ln))

    # ---- declarations
    def declare(self, u, ln):
        m = re.match(r"^(integer|real|complex|double precision|type|logical)\s*(\([^)]*\))?\s*(.*)$", ln)
        base, kind, rest = m.group(1), (m.group(2) or ""), m.group(3)
        if "::" in rest:
            attrs, ents = rest.split("::", 1)
        else:
            attrs, ents = "", rest
            if attrs == "" and ents.startswith(","):
                raise SyntaxError("attributes without '::' in %r" % ln)
        attrs = attrs.strip().lstrip(",")
        kind = kind.replace(" ", "")
        if base == "integer":
            dt = "i"
        elif base == "logical":
            dt = "b"
        elif base == "type":
            dt = "p"
        elif base == "complex":
            dt = "c"
        elif base == "double precision":
            dt = "r"
        else:
            dt = "r" if ("8" in kind or "c_double" in kind) else "r4"
        dim_attr = None
        allocatable = False
        for a in split_top(attrs):
            if a.startswith("dimension"):
                dim_attr = a[a.index("(") + 1:a.rindex(")")]
            if a == "allocatable":
                allocatable = True
        for ent in split_top(ents):
            if not ent:
                continue
            init = None
            depth, cut = 0, None
            for k, ch in enumerate(ent):
                depth += ch == "("
                depth -= ch == ")"
                if ch == "=" and depth == 0:
                    cut = k
                    break
            if cut is not None:
                ent, init = ent[:cut].strip(), ent[cut + 1:].strip()
            m2 = re.match(r"^(\w+)\s*(?:\((.*)\))?$", ent)
            name, dims = m2.group(1), m2.group(2)
            dims = dims if dims is not None else dim_attr
            if dims is not None:
                bounds = []
                for d in split_top(dims):
                    if d == ":":
                        bounds = None
                        break
                    lo, hi = (d.split(":", 1) + [None])[:2] if ":" in d else ("1", d)
                    bounds.append((lo.strip(), hi.strip()))
                u.arrays[name] = (dt, None if (allocatable or bounds is None) else bounds)
            else:
                u.scalars[name] = dt
            if init is not None:
                u.inits.append((name, init))

    NPDT = {"i": "np.int64", "r": "np.float64", "r4": "np.float32", "c": "np.complex128", "b": "np.bool_", "p": "object"}

    def ex(self, u, src):
        p = Parser(tokenize(src), u)
        out = p.expr()
        if not p.done():
            raise SyntaxError("trailing tokens in %r" % src)
        return out

    def assign(self, u, lhs, rhs):
        rhs_py = self.ex(u, rhs)
        m = re.match(r"^(\w+)\s*(\(.*\))?$", lhs.strip())
        if not m:
            raise SyntaxError("bad assignment target %r" % lhs)
        name, idx = m.group(1), m.group(2)
        if name in u.arrays:
            if idx is None:
                return "%s.d[...] = %s" % (u.py(name), rhs_py)
            p = Parser(tokenize(idx[1:-1]), u)
            args = p.arglist()
            return "%s[%s] = %s" % (u.py(name), ", ".join(args) + ("," if len(args) == 1 else ""), rhs_py)
        if idx is not None:
            raise SyntaxError("indexed assignment to undeclared array %r" % lhs)
        dt = u.scalars.get(name)
        if dt is None:
            raise SyntaxError("assignment to undeclared %r" % name)
        conv = {"i": "_int(%s)", "r": "float(%s)", "r4": "_r4(%s)", "c": "complex(%s)", "b": "bool(%s)", "p": "%s"}[dt]
        return "%s = %s" % (u.py(name), conv % rhs_py)

    # ---- one subroutine -> python source
    def translate(self, name):
        u = self.units[name]
        decl_lines, exec_lines = [], []
        for no, ln in u.body:
            if ln.startswith(("use ", "use,", "implicit", "include")):
                continue
            if DECL.match(ln) and not re.match(r"^(real|integer|complex)\s*\(.*\)\s*=", ln) and not exec_lines:
                decl_lines.append(ln)
            else:
                exec_lines.append((no, ln))
        for ln in decl_lines:
            self.declare(u, ln)
        uses_tid = any("omp_get_thread_num" in ln for _, ln in exec_lines)
        out = ["def %s(%s):" % (name, ", ".join(u.py(a) for a in u.args))]
        ind = 1

        def emit(s):
            out.append("    " * ind + s)

        # dummy arrays -> FArr with the declared bounds; dummy scalars -> python scalars
        for a in u.args:
            if a in u.arrays:
                dt, bounds = u.arrays[a]
                lbs = ", ".join("(%s)" % self.ex(u, lo) for lo, _ in bounds)
                shp = ", ".join("(%s) - (%s) + 1" % (self.ex(u, hi), self.ex(u, lo)) for lo, hi in bounds)
                emit("assert isinstance(%s, np.ndarray) and %s.shape == (%s,), ('%s', %s.shape, (%s,))" %
                     (u.py(a), u.py(a), shp, a, u.py(a), shp))
                emit("assert %s.dtype == %s, ('%s', %s.dtype)" % (u.py(a), self.NPDT[dt], a, u.py(a)))
                emit("%s = FArr(%s, (%s,))" % (u.py(a), u.py(a), lbs))
            elif a in u.scalars:
                conv = {"i": "int(%s)", "r": "float(%s)", "r4": "_r4(%s)", "c": "complex(%s)", "b": "bool(%s)", "p": "%s"}[u.scalars[a]]
                emit("%s = %s" % (u.py(a), conv % u.py(a)))
            else:
                raise SyntaxError("%s: dummy argument %s is not declared" % (name, a))
        for a, (dt, bounds) in u.arrays.items():
            if a in u.args:
                continue
            if bounds is None:
                emit("%s = None" % u.py(a))
                continue
            lbs = ", ".join("(%s)" % self.ex(u, lo) for lo, _ in bounds)
            shp = ", ".join("(%s) - (%s) + 1" % (self.ex(u, hi), self.ex(u, lo)) for lo, hi in bounds)
            # locals are undefined until assigned: NaN / a large negative integer makes a read-before-write visible
            fill = {"i": "-2**40", "r": "np.nan", "r4": "np.nan", "c": "complex(np.nan, np.nan)", "b": "False", "p": "None"}[dt]
            emit("%s = FArr(np.full((%s,), %s, dtype=%s, order='F'), (%s,))" % (u.py(a), shp, fill, self.NPDT[dt], lbs))
        for nm, init in u.inits:
            emit(self.assign(u, nm, init))
        emit("_nthreads = 1")
        region = None  # inside an emulated team region
        stack = []
        i = 0
        lines = exec_lines
        while i < len(lines):
            no, ln = lines[i]
            i += 1
            try:
                if ln.startswith("!$omp"):
                    d = ln[5:].strip()
                    if uses_tid and re.match(r"^parallel\b(?!\s+do)", d):
                        priv = re.search(r"private\s*\(([^)]*)\)", d)
                        privs = [x.strip() for x in priv.group(1).split(",")] if priv else []
                        # scalars assigned inside the region and not private are shared
                        j, assigned = i, set()
                        while not re.match(r"^!\$omp\s+end\s+parallel", lines[j][1]):
                            mm = re.match(r"^(\w+)\s*=[^=]", lines[j][1])
                            if mm and mm.group(1) in u.scalars:
                                assigned.add(mm.group(1))
                            mm = re.match(r"^do\s+(\w+)\s*=", lines[j][1])
                            if mm:
                                assigned.add(mm.group(1))
                            j += 1
                        parr = [a for a in privs if a in u.arrays and u.arrays[a][1] is not None]
                        # private arrays: every member of the team works on its own copy (bound through a default
                        # argument, a plain assignment would make the name local before it is read)
                        emit("def _region(_tid, _outer=(%s)):" % "".join(u.py(a) + ", " for a in parr))
                        ind += 1
                        shared = sorted(a for a in assigned if a not in privs)
                        if shared:
                            emit("nonlocal " + ", ".join(u.py(a) for a in shared))
                        for k, a in enumerate(parr):
                            emit("%s = _outer[%d].copy()" % (u.py(a), k))
                        region = True
                        continue
                    if region and re.match(r"^end\s+parallel\b(?!\s+do)", d):
                        emit("yield")
                        ind -= 1
                        emit("_run_team(_region, _nthreads)")
                        region = None
                        continue
                    if d.startswith("barrier"):
                        if region:
                            emit("yield")
                        continue
                    continue  # every other directive: one thread executes the region in order
                m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", ln)
                if m:
                    parts = split_top(m.group(2))
                    a, b = self.ex(u, parts[0]), self.ex(u, parts[1])
                    st = self.ex(u, parts[2]) if len(parts) > 2 else "1"
                    v = u.py(m.group(1))
                    emit("for %s in (range(int(%s), int(%s) + 1, int(%s)) if int(%s) > 0 else range(int(%s), int(%s) - 1, int(%s))):"
                         % (v, a, b, st, st, a, b, st))
                    ind += 1
                    stack.append("do")
                    continue
                if re.match(r"^end\s*do$", ln):
                    assert stack.pop() == "do"
                    ind -= 1
                    continue
                m = re.match(r"^if\s*\((.*)\)\s*then$", ln)
                if m:
                    emit("if %s:" % self.ex(u, m.group(1)))
                    ind += 1
                    stack.append("if")
                    emit("pass")
                    continue
                m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", ln)
                if m:
                    ind -= 1
                    emit("elif %s:" % self.ex(u, m.group(1)))
                    ind += 1
                    emit("pass")
                    continue
                if ln == "else":
                    ind -= 1
                    emit("else:")
                    ind += 1
                    emit("pass")
                    continue
                if re.match(r"^end\s*if$", ln):
                    assert stack.pop() == "if"
                    ind -= 1
                    continue
                m = re.match(r"^if\s*\(", ln)
                if m:  # one-line if: find the matching parenthesis
                    depth, k = 0, ln.index("(")
                    for k in range(ln.index("("), len(ln)):
                        depth += ln[k] == "("
                        depth -= ln[k] == ")"
                        if depth == 0:
                            break
                    cond, stmt = ln[ln.index("(") + 1:k], ln[k + 1:].strip()
                    emit("if %s:" % self.ex(u, cond))
                    ind += 1
                    emit(self.simple(u, stmt))
                    ind -= 1
                    continue
                emit(self.simple(u, ln))
            except Exception as e:
                raise SyntaxError("%s:%d: %s   [%s]" % (os.path.basename(self.path), no, e, ln)) from e
        assert not stack, (name, stack)
        emit("return {%s}" % ", ".join("'%s': %s" % (a, u.py(a)) for a in u.args if a in u.scalars))
        return "\n".join(out) + "\n"

    def simple(self, u, ln):
        if ln == "cycle":
            return "continue"
        if ln == "exit":
            return "break"
        if ln == "return":
            raise SyntaxError("early return is not supported")
        m = re.match(r"^call\s+(\w+)\s*\((.*)\)$", ln)
        if m:
            fn, args = m.group(1), split_top(m.group(2))
            if fn == "omp_set_num_threads":
                return "_nthreads = %s" % self.ex(u, args[0])
            if fn in ("fftw_execute_dft", "dfftw_execute_dft"):
                return "_fftw_exec(%s, %s, %s)" % (u.py(args[0]), u.py(args[1]), u.py(args[2]))
            if fn in ("fftw_destroy_plan", "dfftw_destroy_plan"):
                return "pass"
            raise SyntaxError("call to %s is not supported" % fn)
        m = re.match(r"^allocate\s*\((.*)\)$", ln)
        if m:
            outs = []
            for ent in split_top(m.group(1)):
                m2 = re.match(r"^(\w+)\s*\((.*)\)$", ent)
                nm, dims = m2.group(1), split_top(m2.group(2))
                dt = u.arrays[nm][0]
                los, shp = [], []
                for d in dims:
                    lo, hi = (d.split(":", 1)) if ":" in d else ("1", d)
                    los.append("(%s)" % self.ex(u, lo))
                    shp.append("(%s) - (%s) + 1" % (self.ex(u, hi), self.ex(u, lo)))
                fill = {"i": "-2**40", "r": "np.nan", "r4": "np.nan", "c": "complex(np.nan, np.nan)"}[dt]
                outs.append("%s = FArr(np.full((%s,), %s, dtype=%s, order='F'), (%s,))" %
                            (u.py(nm), ", ".join(shp), fill, self.NPDT[dt], ", ".join(los)))
            return "; ".join(outs)
        if re.match(r"^deallocate\s*\(", ln):
            return "pass"
        # assignment: split at the first top-level '=' that is not part of ==, /=, <=, >=
        depth = 0
        for k, ch in enumerate(ln):
            depth += ch == "("
            depth -= ch == ")"
            if ch == "=" and depth == 0 and ln[k + 1:k + 2] != "=" and ln[k - 1] not in "=/<>":
                return self.assign(u, ln[:k], ln[k + 1:])
        raise SyntaxError("statement not understood")


class F90Module:
    """the subroutines of a set of .f90 files as python callables (full Fortran argument lists, hidden dimensions
    included; arrays are modified in place, scalar dummies come back in a dict)"""

    def __init__(self, paths):
        self.src, self.fn, self.warnings = {}, {}, {}
        for p in paths:
            t = Translator(p)
            for name in t.units:
                try:
                    code = t.translate(name)
                except SyntaxError as e:
                    self.src[name] = e
                    continue
                self.src[name] = code
                self.warnings[name] = t.units[name].warnings
                ns = dict(RUNTIME)
                exec(compile(code, "<f90:%s:%s>" % (os.path.basename(p), name), "exec"), ns)  # noqa: S102
                self.fn[name] = ns[name]

    def __getattr__(self, name):
        try:
            return self.fn[name]
        except KeyError:
            err = self.src.get(name)
            raise AttributeError("%s: %s" % (name, err if err is not None else "no such subroutine"))

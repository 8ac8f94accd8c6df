#!/usr/bin/env python
"""f90toc.py -- machine-translates the hot-path subroutines of the reference into C.

TEST INFRASTRUCTURE ONLY (see oracle/d3q19_oracle.h).  The reference is Fortran 90 + MPI and
this image has neither a Fortran compiler nor MPI, so the reference cannot be built as it is.
This script reads the reference's own sources WHERE THEY LIE (default
/root/reference/Channel-Flow) and emits an equivalent C translation unit into oracle/_ref/
(git-ignored; nothing of the reference is copied into the repository):

    var_inc.f90   every module variable  -> a field of `struct ref_state` (one per MPI rank)
    para.f90      para (up to `rhoepsl = ...`, i.e. without the directory/particle set-up),
                  allocarray
    initial.f90   initpop, initvel
    collision.f90 collision_MRT, collisionExchnge, macrovar, rhoupdat, avedensity, FORCING, FORCINGP
    saveload.f90  vortcalc, exchng8, statistc, statistc2, diag (values written to file units are
                  captured: ref_capture)
    saveload.f90  outputflow, outputuy, outputpress, probe (what main.f90 calls every nflowout steps and at the end)
    main.f90      PROGRAM main itself -> ref_main (MPI_WTIME is a call counter; constructMPItypes is skipped -- derived
                  MPI types are not used by the mini-MPI -- and any other call outside the translated set
                  aborts instead of being passed over)

With -DREF_DROPIN the subroutines of collision.f90 are left out and `shim_translated.c` (the translation of
THIS repository's fortran/collision_b200.f90 by oracle/shim2c.py) is included in their place: the
reference's driver linked against the drop-in instead of its own collision.f90 (_ref/libref_b200.so).

The translation is statement by statement: every expression keeps the Fortran evaluation
order (left to right for equal precedence; `**` by repeated multiplication), `real` is double
(the reference is built with -r8, Makefile:29), `integer` is int, array sections become loop
nests, `goto`/labels stay, automatic arrays are heap-allocated for the call, and MPI calls go
to the in-process mini-MPI of oracle/ref_runtime.h where each rank is a thread.  Compiled with
-O2 -ffp-contract=off the result is the IEEE-754 evaluation of the reference text; it is what
the hand-written restatement (oracle/d3q19_oracle.c) is pinned against, bit for bit
(tests/test_oracle_ref.py), and what generated tests/golden/*.npz.

Run-time configuration the reference fixes at compile time or hard-codes (grid size nx7/nx/ny/nz,
`laminarFlow`, `nprocY`, physical scalars set in `para`) can be overridden by name through
ref_set_override(); an override replaces the right-hand side of the assignment (or the
parameter's initialiser) and nothing else.

Usage: python f90toc.py --ref /root/reference/Channel-Flow --out oracle/_ref/ref_translated.c
"""
import argparse
import os
import re
import sys

# ------------------------------------------------------------------------------------------------
# source reading
# ------------------------------------------------------------------------------------------------


def strip_comment(line):
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


def lower_outside_strings(s):
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


def split_concat(s):
    """pieces of a character expression a // b // c (top level, outside strings)"""
    out, cur, depth, q, i = [], "", 0, None, 0
    while i < len(s):
        ch = s[i]
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch; cur += ch
        elif ch == "(":
            depth += 1; cur += ch
        elif ch == ")":
            depth -= 1; cur += ch
        elif ch == "/" and s[i:i + 2] == "//" and depth == 0:
            out.append(cur); cur = ""; i += 1
        else:
            cur += ch
        i += 1
    out.append(cur)
    return out


def logical_lines(path):
    """[(first_line_number, text)] with comments removed, continuations joined, lower-cased."""
    res, cur, start = [], "", None
    for no, raw in enumerate(open(path, encoding="latin-1").read().splitlines(), 1):
        line = strip_comment(raw.replace("\t", " "))
        if not line.strip():
            continue
        body = line.strip()
        if cur:
            if body.startswith("&"):
                body = body[1:].lstrip()
            cur += " " + body
        else:
            cur, start = body, no
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        res.append((start, lower_outside_strings(cur)))
        cur = ""
    return res


# ------------------------------------------------------------------------------------------------
# expressions
# ------------------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
    (?P<dotop>\.(?:and|or|not|lt|le|gt|ge|eq|ne|true|false)\.)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?)
  | (?P<id>[a-z_][a-z0-9_]*)
  | (?P<op>\*\*|\(/|/\)|==|/=|<=|>=|[-+*/(),:<>=%])
  | (?P<ws>\s+)
""", re.X)


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = TOKEN_RE.match(s, pos)
        if not m:
            raise SyntaxError("cannot tokenize %r at %d" % (s, pos))
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        text = m.group()
        if kind == "num":
            # "1.lt." style collisions: a number must not swallow the dot of a dot-operator
            if text.endswith(".") and re.match(r"(?:and|or|not|lt|le|gt|ge|eq|ne)\.", s[pos:]):
                text = text[:-1]
                pos -= 1
        toks.append((kind, text))
    return toks


class Num:
    def __init__(self, text):
        self.text = text
        self.is_real = bool(re.search(r"[.ed]", text))


class Var:
    def __init__(self, name):
        self.name = name


class Index:           # array element / section reference or function call
    def __init__(self, name, args):
        self.name, self.args = name, args


class Range:
    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi


class Kw:
    def __init__(self, key, val):
        self.key, self.val = key, val


class Un:
    def __init__(self, op, e):
        self.op, self.e = op, e


class Bin:
    def __init__(self, op, l, r):
        self.op, self.l, self.r = op, l, r


class Cons:
    def __init__(self, items):
        self.items = items


class Logical:
    def __init__(self, v):
        self.v = v


REL = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "!=",
       "<": "<", "<=": "<=", ">": ">", ">=": ">=", "==": "==", "/=": "!="}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i][1] if self.i < len(self.t) else None

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, text):
        if self.peek() == text:
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            raise SyntaxError("expected %r, got %r in %r" % (text, self.peek(), self.t))

    def done(self):
        return self.i >= len(self.t)

    # .or. < .and. < .not. < relational < additive (with leading sign) < multiplicative < power
    def expr(self):
        e = self.and_()
        while self.peek() == ".or.":
            self.next()
            e = Bin("||", e, self.and_())
        return e

    def and_(self):
        e = self.not_()
        while self.peek() == ".and.":
            self.next()
            e = Bin("&&", e, self.not_())
        return e

    def not_(self):
        if self.peek() == ".not.":
            self.next()
            return Un("!", self.not_())
        return self.rel()

    def rel(self):
        e = self.add()
        if self.peek() in REL:
            op = REL[self.next()[1]]
            e = Bin(op, e, self.add())
        return e

    def add(self):
        if self.peek() in ("+", "-"):
            op = self.next()[1]
            e = self.mul()
            if op == "-":
                e = Un("-", e)
        else:
            e = self.mul()
        while self.peek() in ("+", "-"):
            op = self.next()[1]
            e = Bin(op, e, self.mul())
        return e

    def mul(self):
        e = self.pow_()
        while self.peek() in ("*", "/"):
            op = self.next()[1]
            e = Bin(op, e, self.pow_())
        return e

    def pow_(self):
        e = self.primary()
        if self.peek() == "**":
            self.next()
            if self.peek() in ("+", "-"):
                sgn = self.next()[1]
                r = self.pow_()
                r = Un("-", r) if sgn == "-" else r
            else:
                r = self.pow_()
            e = Bin("**", e, r)
        return e

    def primary(self):
        kind, text = self.next()
        if kind == "num":
            return Num(text)
        if kind == "dotop":
            if text == ".true.":
                return Logical(1)
            if text == ".false.":
                return Logical(0)
            raise SyntaxError("unexpected %s" % text)
        if kind == "id":
            if self.peek() == "(":
                self.next()
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.arg())
                        if self.accept(")"):
                            break
                        self.expect(",")
                return Index(text, args)
            return Var(text)
        if text == "(":
            e = self.expr()
            self.expect(")")
            return Un("()", e)
        if text == "(/":
            items = []
            while True:
                items.append(self.expr())
                if self.accept("/)"):
                    break
                self.expect(",")
            return Cons(items)
        raise SyntaxError("unexpected token %r in %r" % (text, self.t))

    def arg(self):
        # keyword argument (mask = ...)?
        if (self.i + 1 < len(self.t) and self.t[self.i][0] == "id" and self.t[self.i + 1][1] == "="
                and (self.i + 2 >= len(self.t) or self.t[self.i + 2][1] != "=")):
            key = self.next()[1]
            self.next()
            return Kw(key, self.expr())
        lo = None
        if self.peek() != ":":
            lo = self.expr()
            if self.peek() != ":":
                return lo
        self.expect(":")
        hi = None
        if self.peek() not in (",", ")"):
            hi = self.expr()
        return Range(lo, hi)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if not p.done():
        raise SyntaxError("trailing tokens in %r" % s)
    return e


def split_top(s, sep=","):
    """split at top-level separators (outside parentheses and strings)"""
    out, depth, cur, q = [], 0, "", None
    i = 0
    while i < len(s):
        ch = s[i]
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == "(":
            depth += 1
            cur += ch
        elif ch == ")":
            depth -= 1
            cur += ch
        elif ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip():
        out.append(cur.strip())
    return out


def matching_paren(s, start):
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise SyntaxError("unbalanced parentheses in %r" % s)


# ------------------------------------------------------------------------------------------------
# symbols
# ------------------------------------------------------------------------------------------------
class Sym:
    def __init__(self, name, typ, dims=None, param=None, allocatable=False, scope="module", dummy=False):
        self.name, self.typ = name, typ          # typ: real | int | logical | other
        self.dims = dims                         # None (scalar) or list of (lo_expr_or_None, hi_expr_or_None)
        self.param = param                       # initialiser expression text for parameters
        self.allocatable = allocatable
        self.scope = scope                       # module | local
        self.dummy = dummy

    @property
    def rank(self):
        return len(self.dims) if self.dims else 0


TYPE_RE = re.compile(r"^(real|integer|logical|double precision|character|type)\b(\s*\([^)]*\))?")


def parse_decl(text):
    """-> (typ, attrs, [(name, dims_text_or_None, init_text_or_None)]) or None"""
    m = TYPE_RE.match(text)
    if not m:
        return None
    base = m.group(1)
    if base in ("character", "type") and not text.startswith("type("):
        if base == "type":
            return None
    typ = {"real": "real", "double precision": "real", "integer": "int", "logical": "logical"}.get(base, "other")
    rest = text[m.end():]
    if base == "type" and text.startswith("type("):
        rest = text[matching_paren(text, 4) + 1:]
        typ = "other"
    if m.group(2) and base == "integer" and "8" in m.group(2):
        typ = "other"                           # integer(kind=8) FFTW plans, unused on the path
    attrs = {}
    if "::" in rest:
        a, names = rest.split("::", 1)
        for item in split_top(a.strip().lstrip(",")):
            item = item.strip()
            if item.startswith("dimension"):
                attrs["dimension"] = item[item.index("(") + 1: matching_paren(item, item.index("("))]
            elif item:
                attrs[item] = True
    else:
        names = rest
    ents = []
    for ent in split_top(names):
        init = None
        if "=" in ent and typ != "other":
            # parameter initialiser (not ==)
            k = ent.index("=")
            ent, init = ent[:k].strip(), ent[k + 1:].strip()
        dims = None
        if "(" in ent:
            k = ent.index("(")
            dims = ent[k + 1: matching_paren(ent, k)]
            ent = ent[:k].strip()
        if "*" in ent:                          # character*n
            ent = ent.split("*")[0].strip()
        ents.append((ent.strip(), dims, init))
    return typ, attrs, ents


def parse_dims(dims_text):
    dims = []
    for d in split_top(dims_text):
        if d == ":":
            dims.append((None, None))
        elif ":" in d and not d.startswith("("):
            lo, hi = d.split(":", 1)
            dims.append((lo.strip(), hi.strip()))
        else:
            dims.append(("1", d.strip()))
    return dims


# ------------------------------------------------------------------------------------------------
# translation
# ------------------------------------------------------------------------------------------------
INTRINSIC_REAL = {"exp": "exp", "sin": "sin", "cos": "cos", "alog": "log", "log": "log", "dlog": "log",
                  "atan": "atan", "sqrt": "sqrt", "dsqrt": "sqrt", "dcos": "cos", "dsin": "sin", "dexp": "exp"}
MPI_CONST = {"mpi_real8": "REF_MPI_REAL8", "mpi_integer": "REF_MPI_INTEGER", "mpi_sum": "REF_MPI_SUM",
             "mpi_max": "REF_MPI_MAX", "mpi_min": "REF_MPI_MIN", "mpi_comm_world": "0", "mpi_status_size": "4",
             "mpi_byte": "REF_MPI_BYTE"}


class Translator:
    def __init__(self, ref_dir):
        self.ref = ref_dir
        self.module = {}           # name -> Sym
        self.module_order = []
        self.local = {}
        self.out = []
        self.subs = {}             # name -> [dummy names]
        self.cur_sub = None
        self.tmp_id = 0
        self.parse_module()

    # ---- var_inc.f90 ---------------------------------------------------------------------------
    def parse_module(self):
        lines = logical_lines(os.path.join(self.ref, "var_inc.f90"))
        in_type = False
        for no, text in lines:
            if re.match(r"^type\s+[a-z_]", text) and not text.startswith("type("):
                in_type = True
                continue
            if in_type:
                if re.match(r"^end\s*type", text):
                    in_type = False
                continue
            d = parse_decl(text)
            if not d:
                continue
            typ, attrs, ents = d
            for name, dims, init in ents:
                dtext = dims if dims is not None else attrs.get("dimension")
                sym = Sym(name, typ, parse_dims(dtext) if dtext else None,
                          param=init if "parameter" in attrs else None,
                          allocatable="allocatable" in attrs)
                self.module[name] = sym
                self.module_order.append(name)

    # ---- lookups ---------------------------------------------------------------------------------
    def sym(self, name):
        return self.local.get(name) or self.module.get(name)

    def c_name(self, sym):
        return "S->%s" % sym.name if sym.scope == "module" else sym.name

    # ---- types -----------------------------------------------------------------------------------
    def typeof(self, e):
        if isinstance(e, Num):
            return "real" if e.is_real else "int"
        if isinstance(e, Logical):
            return "logical"
        if isinstance(e, Var):
            s = self.sym(e.name)
            if not s:
                if e.name in MPI_CONST:
                    return "int"
                raise KeyError("unknown variable %r in %s" % (e.name, self.cur_sub))
            return s.typ
        if isinstance(e, Index):
            s = self.sym(e.name)
            if s and s.dims:
                return s.typ
            n = e.name
            if n in ("real", "dfloat", "dble", "float") or n in INTRINSIC_REAL:
                return "real"
            if n in ("int", "nint", "count", "size"):
                return "int"
            if n == "mpi_wtime":
                return "real"
            if n in ("mod", "abs", "max", "min", "sum", "maxval", "minval", "sign"):
                ts = [self.typeof(a) for a in e.args if not isinstance(a, (Kw, Range))]
                return "real" if "real" in ts else "int"
            raise KeyError("unknown function or array %r in %s" % (n, self.cur_sub))
        if isinstance(e, Un):
            return "logical" if e.op == "!" else self.typeof(e.e)
        if isinstance(e, Bin):
            if e.op in ("||", "&&", "<", "<=", ">", ">=", "==", "!="):
                return "logical"
            tl, tr = self.typeof(e.l), self.typeof(e.r)
            if e.op == "**":
                return tl
            return "real" if "real" in (tl, tr) else "int"
        if isinstance(e, Cons):
            return self.typeof(e.items[0])
        raise TypeError(e)

    # ---- expression -> C -------------------------------------------------------------------------
    def num_c(self, n):
        t = n.text.replace("d", "e")
        if n.is_real:
            if re.fullmatch(r"\d+\.", t):
                t += "0"
            t = re.sub(r"^(\d+)\.e", r"\1.0e", t)
            if not re.search(r"[.e]", t):
                t += ".0"
            return t
        return t

    def elem(self, sym, idx_c):
        """C lvalue of an array element; idx_c are C index expressions (Fortran index values)."""
        A = self.c_name(sym)
        acc = "REF_R" if sym.typ == "real" else "REF_I"
        if len(idx_c) != sym.rank:
            raise SyntaxError("rank mismatch for %s: %d subscripts" % (sym.name, len(idx_c)))
        return "%s%d(%s, %s)" % (acc, sym.rank, A, ", ".join("(%s)" % i for i in idx_c))

    def lo_c(self, sym, d):
        return "%s.lo[%d]" % (self.c_name(sym), d)

    def hi_c(self, sym, d):
        return "(%s.lo[%d] + %s.n[%d] - 1)" % (self.c_name(sym), d, self.c_name(sym), d)

    def cx(self, e, sect=None):
        """C text of expression e.  sect: list of C loop-variable names that replace, in order,
        the section dimensions of every array-valued reference inside e."""
        if isinstance(e, Num):
            return self.num_c(e)
        if isinstance(e, Logical):
            return str(e.v)
        if isinstance(e, Var):
            if e.name in MPI_CONST:
                return MPI_CONST[e.name]
            s = self.sym(e.name)
            if not s:
                raise KeyError("unknown variable %r in %s" % (e.name, self.cur_sub))
            if s.dims:
                if sect is None:
                    raise SyntaxError("whole-array reference %s outside an array statement (%s)" % (e.name, self.cur_sub))
                if len(sect) != s.rank:
                    raise SyntaxError("non-conformable whole array %s" % e.name)
                return self.elem(s, ["%s + %s" % (self.lo_c(s, d), sect[d]) for d in range(s.rank)])
            return self.c_name(s)
        if isinstance(e, Index):
            s = self.sym(e.name)
            if s and s.dims:
                idx, k = [], 0
                for d, a in enumerate(e.args):
                    if isinstance(a, Range):
                        if sect is None:
                            raise SyntaxError("array section of %s outside an array statement (%s)" % (e.name, self.cur_sub))
                        lo = self.cx(a.lo) if a.lo is not None else self.lo_c(s, d)
                        idx.append("%s + %s" % (lo, sect[k]))
                        k += 1
                    else:
                        idx.append(self.cx(a))
                if sect is not None and k not in (0, len(sect)):
                    raise SyntaxError("non-conformable section of %s" % e.name)
                return self.elem(s, idx)
            return self.call_intrinsic(e, sect)
        if isinstance(e, Un):
            if e.op == "()":
                return "(%s)" % self.cx(e.e, sect)
            return "(%s(%s))" % (e.op, self.cx(e.e, sect))
        if isinstance(e, Bin):
            if e.op == "**":
                if isinstance(e.r, Num) and not e.r.is_real:
                    fn = "ref_powi_d" if self.typeof(e.l) == "real" else "ref_powi_i"
                    return "%s(%s, %s)" % (fn, self.cx(e.l, sect), e.r.text)
                return "pow(%s, %s)" % (self.cx(e.l, sect), self.cx(e.r, sect))
            return "(%s %s %s)" % (self.cx(e.l, sect), e.op, self.cx(e.r, sect))
        raise TypeError("cannot translate %r" % e)

    def call_intrinsic(self, e, sect):
        n = e.name
        a = [self.cx(x, sect) for x in e.args if not isinstance(x, (Kw, Range))]
        if n in ("real", "dfloat", "dble", "float"):
            return "((double)(%s))" % a[0]
        if n == "mpi_wtime":
            return "ref_mpi_wtime(S)"          # a counter: 0 unless the test overrides "wtime_tick" (main.f90:197-207)
        if n == "int":
            return "((int)(%s))" % a[0]
        if n in INTRINSIC_REAL:
            return "%s(%s)" % (INTRINSIC_REAL[n], a[0])
        if n == "mod":
            if self.typeof(e) == "int":
                return "((%s) %% (%s))" % (a[0], a[1])
            return "fmod(%s, %s)" % (a[0], a[1])
        if n == "abs":
            return ("fabs(%s)" if self.typeof(e) == "real" else "abs(%s)") % a[0]
        if n in ("max", "min"):
            fn = ("ref_%s_d" if self.typeof(e) == "real" else "ref_%s_i") % n
            r = a[0]
            for x in a[1:]:
                r = "%s(%s, %s)" % (fn, r, x)
            return r
        raise KeyError("unsupported intrinsic %r in %s" % (n, self.cur_sub))

    # ---- sections ---------------------------------------------------------------------------------
    def section_shape(self, e):
        """[(count_c)] of the first array-valued reference found in e (depth first), or None."""
        if isinstance(e, Var):
            s = self.sym(e.name)
            if s and s.dims:
                return ["%s.n[%d]" % (self.c_name(s), d) for d in range(s.rank)]
            return None
        if isinstance(e, Index):
            s = self.sym(e.name)
            if s and s.dims:
                shp = []
                for d, a in enumerate(e.args):
                    if isinstance(a, Range):
                        lo = self.cx(a.lo) if a.lo is not None else self.lo_c(s, d)
                        hi = self.cx(a.hi) if a.hi is not None else self.hi_c(s, d)
                        shp.append("((%s) - (%s) + 1)" % (hi, lo))
                return shp or None
            for a in e.args:
                x = a.val if isinstance(a, Kw) else a
                if isinstance(x, Range):
                    continue
                r = self.section_shape(x)
                if r:
                    return r
            return None
        if isinstance(e, Un):
            return self.section_shape(e.e)
        if isinstance(e, Bin):
            return self.section_shape(e.l) or self.section_shape(e.r)
        return None

    def loops(self, shape):
        """-> (open_text, close_text, loop var names); first dimension innermost (column major)"""
        self.tmp_id += 1
        names = ["s%d_%d" % (self.tmp_id, d) for d in range(len(shape))]
        op = ""
        for d in reversed(range(len(shape))):
            op += "for (int %s = 0; %s < %s; ++%s) " % (names[d], names[d], shape[d], names[d])
        return op + "{ ", " }", names

    # ---- statements -------------------------------------------------------------------------------
    def emit(self, s):
        self.out.append(s)

    def assignment(self, lhs_text, rhs_text, override_ok=False):
        lhs, rhs = parse_expr(lhs_text), parse_expr(rhs_text)
        shape = self.section_shape(lhs)
        if shape:                                               # array statement
            if isinstance(rhs, Cons):
                s = self.sym(lhs.name)
                for k, it in enumerate(rhs.items):
                    self.emit("%s = %s;" % (self.elem(s, ["%s + %d" % (self.lo_c(s, 0), k)]), self.cx(it)))
                return
            op, cl, names = self.loops(shape)
            self.emit("%s%s = %s;%s" % (op, self.cx(lhs, names), self.cx(rhs, names), cl))
            return
        # scalar on the left: reductions on the right?
        if isinstance(rhs, Index) and rhs.name in ("count", "sum", "maxval", "minval") \
                and not (self.sym(rhs.name) and self.sym(rhs.name).dims):
            self.reduction(lhs, rhs)
            return
        r = self.cx(rhs)
        if override_ok and isinstance(lhs, Var) and self.sym(lhs.name).scope in ("module", "local"):
            t = self.sym(lhs.name).typ
            fn = "ref_override_d" if t == "real" else "ref_override_i"
            r = '%s(S, "%s", %s)' % (fn, lhs.name, r)
        self.emit("%s = %s;" % (self.cx(lhs), r))

    def reduction(self, lhs, rhs):
        pos = [a for a in rhs.args if not isinstance(a, Kw)]
        kws = {a.key: a.val for a in rhs.args if isinstance(a, Kw)}
        shape = self.section_shape(pos[0])
        op, cl, names = self.loops(shape)
        acc = self.cx(lhs)
        if rhs.name == "count":
            self.emit("%s = 0; %sif (%s) %s += 1;%s" % (acc, op, self.cx(pos[0], names), acc, cl))
        elif rhs.name in ("maxval", "minval"):
            cond = "if (%s) " % self.cx(kws["mask"], names) if "mask" in kws else ""
            self.tmp_id += 1
            t = "red%d" % self.tmp_id
            init, cmp_ = ("-HUGE_VAL", ">") if rhs.name == "maxval" else ("HUGE_VAL", "<")
            self.emit("{ double %s = %s; %s%sif (%s %s %s) %s = %s;%s %s = %s; }"
                      % (t, init, op, cond, self.cx(pos[0], names), cmp_, t, t, self.cx(pos[0], names), cl, acc, t))
        else:
            cond = "if (%s) " % self.cx(kws["mask"], names) if "mask" in kws else ""
            # accumulate in a temporary like the intrinsic does, then assign
            self.tmp_id += 1
            t = "red%d" % self.tmp_id
            self.emit("{ double %s = 0.0; %s%s%s += %s;%s %s = %s; }" % (t, op, cond, t, self.cx(pos[0], names), cl, acc, t))

    def call(self, text):
        m = re.match(r"call\s+([a-z_0-9]+)\s*(\((.*)\))?\s*$", text)
        name, args = m.group(1), split_top(m.group(3)) if m.group(3) else []
        if name.startswith("mpi_"):
            cargs = []
            for a in args:
                e = parse_expr(a)
                if isinstance(e, Var) and e.name in MPI_CONST:
                    cargs.append(MPI_CONST[e.name])
                    continue
                cargs.append(self.by_ref(e))
            self.emit("ref_%s(S, %s);" % (name, ", ".join(cargs)))
            return
        if name in self.wanted:
            cargs, after = [], []
            for a in args:
                e = parse_expr(a)
                sec = isinstance(e, Index) and self.sym(e.name) and self.sym(e.name).dims \
                    and any(isinstance(x, Range) for x in e.args)
                if not sec:
                    cargs.append(self.by_ref(e))
                    continue
                # array-section actual argument (e.g. ux(:,1,:)): copy-in / copy-out through a contiguous
                # temporary, which is what a Fortran compiler does for a non-contiguous section
                s0 = self.sym(e.name)
                shape = self.section_shape(e)
                self.tmp_id += 1
                t = "sec%d" % self.tmp_id
                kind, acc = ("REF_KIND_R", "REF_R") if s0.typ == "real" else ("REF_KIND_I", "REF_I")
                self.emit("ref_arr %s = ref_alloc_auto(%s, %d, (int[]){%s}, (int[]){%s});"
                          % (t, kind, len(shape), ", ".join("1" for _ in shape), ", ".join(shape)))
                op, cl, names = self.loops(shape)
                tel = "%s%d(%s, %s)" % (acc, len(shape), t, ", ".join("1 + %s" % n for n in names))
                self.emit("%s%s = %s;%s" % (op, tel, self.cx(e, names), cl))
                op2, cl2, names2 = self.loops(shape)
                tel2 = "%s%d(%s, %s)" % (acc, len(shape), t, ", ".join("1 + %s" % n for n in names2))
                after.append("%s%s = %s;%s ref_free(&%s);" % (op2, self.cx(e, names2), tel2, cl2, t))
                cargs.append("%s.p" % t)
            self.emit("ref_%s(%s);" % (name, ", ".join(["S"] + cargs)))
            for a in after:
                self.emit(a)
            return
        if self.cur_sub == "main" and name != "constructmpitypes":
            self.emit('ref_untranslated(S, "%s");' % name)
            return
        self.emit("/* call %s skipped (outside the translated path) */;" % name)

    def by_ref(self, e):
        """Fortran passes by reference: arrays -> data pointer, scalars -> address, rvalues -> temporaries."""
        if isinstance(e, Var):
            s = self.sym(e.name)
            if s and s.dims:
                return "%s.p" % self.c_name(s)
            if s:
                return "&%s" % self.c_name(s)
        if isinstance(e, Index):
            s = self.sym(e.name)
            if s and s.dims and not any(isinstance(a, Range) for a in e.args):
                return "&%s" % self.cx(e)
        t = self.typeof(e)
        return "&(%s){%s}" % ("double" if t == "real" else "int", self.cx(e))

    def declare_local(self, text, dummies):
        d = parse_decl(text)
        if not d:
            return False
        typ, attrs, ents = d
        for name, dims, init in ents:
            dtext = dims if dims is not None else attrs.get("dimension")
            sym = Sym(name, typ, parse_dims(dtext) if dtext else None, scope="local", dummy=name in dummies)
            self.local[name] = sym
        return True

    def emit_local_decls(self):
        """C declarations of this subroutine's locals; automatic arrays are allocated on entry."""
        frees = []
        for s in self.local.values():
            ctype = {"real": "double", "int": "int", "logical": "int"}.get(s.typ)
            if ctype is None:
                continue
            if not s.dims:
                if s.dummy:
                    self.emit("#define %s (*(%s *)%s_arg)" % (s.name, ctype, s.name))
                else:
                    self.emit("%s %s = 0; (void)%s;" % (ctype, s.name, s.name))
                continue
            if any(lo is None or hi is None for lo, hi in s.dims):
                # a deferred-shape (allocatable) local: it is allocated, if at all, beyond the translated part
                self.emit("ref_arr %s; memset(&%s, 0, sizeof %s); (void)%s;" % (s.name, s.name, s.name, s.name))
                continue
            los = ", ".join("(%s)" % self.cx(parse_expr(lo)) for lo, hi in s.dims)
            his = ", ".join("(%s)" % self.cx(parse_expr(hi)) for lo, hi in s.dims)
            kind = "REF_KIND_R" if s.typ == "real" else "REF_KIND_I"
            if s.dummy:
                self.emit("ref_arr %s = ref_view(%s_arg, %s, %d, (int[]){%s}, (int[]){%s});"
                          % (s.name, s.name, kind, s.rank, los, his))
            else:
                self.emit("ref_arr %s = ref_alloc_auto(%s, %d, (int[]){%s}, (int[]){%s});" % (s.name, kind, s.rank, los, his))
                frees.append("ref_free(&%s);" % s.name)
        return frees

    def translate_sub(self, name, lines, stop_after=None, override_assignments=False, capture_local=None):
        """lines: logical lines of one subroutine (from `subroutine` to `end subroutine`).
        stop_after: a regex, or (regex, n): translation stops after the n-th statement that matches.
        capture_local: (array name, unit): before returning, every element of that LOCAL array goes to the capture
        buffer under `unit` (first dimension fastest) -- how a result that the reference keeps in an automatic array
        (sijstat00's Sij2) becomes visible to the tests."""
        self.local, self.cur_sub = {}, name
        head = lines[0][1]
        m = re.match(r"subroutine\s+([a-z_0-9]+)\s*(\(([^)]*)\))?", head)
        dummies = [a.strip() for a in m.group(3).split(",")] if m.group(3) else []
        self.subs[name] = dummies
        body = lines[1:]
        # declarations first
        k = 0
        decl_lines = []
        while k < len(body):
            t = body[k][1]
            if t.startswith("use ") or t.startswith("implicit ") or t.startswith("external "):
                k += 1
                continue
            if TYPE_RE.match(t):
                decl_lines.append(t)
                k += 1
                continue
            break
        for t in decl_lines:
            self.declare_local(t, dummies)
        sig = ", ".join(["ref_state *S"] + ["void *%s_arg" % d for d in dummies])
        self.emit("\nvoid ref_%s(%s)\n{" % (name, sig))
        frees = self.emit_local_decls()
        stop_re, stop_n = (stop_after if isinstance(stop_after, tuple) else (stop_after, 1))
        for no, t in body[k:]:
            if re.match(r"^end\s*subroutine", t) or t == "end":
                break
            try:
                self.statement(t, override_assignments)
            except Exception as ex:
                raise type(ex)("%s (line %d of %s: %r)" % (ex, no, name, t)) from ex
            if stop_re and re.match(stop_re, t):
                stop_n -= 1
                if stop_n == 0:
                    break
        self.emit("ref_end: ;")
        if capture_local:
            e = parse_expr(capture_local[0])
            op, cl, names = self.loops(self.section_shape(e))
            self.emit("%sref_capture(S, %d, (double)(%s));%s" % (op, capture_local[1], self.cx(e, names), cl))
        for f in frees:
            self.emit(f)
        for s in self.local.values():
            if s.dummy and not s.dims and s.typ in ("real", "int", "logical"):
                self.emit("#undef %s" % s.name)
        self.emit("}")

    def statement(self, t, ov=False):
        # label
        m = re.match(r"^(\d+)\s+(.*)$", t)
        if m:
            self.emit("L%s: ;" % m.group(1))
            t = m.group(2).strip()
            if t == "continue":
                return
        if t == "continue":
            return
        m = re.match(r"^write\s*\(\s*(\d+)\s*\)\s*(.*)$", t)
        if m:
            # UNFORMATTED sequential output (checkpoint files, saveload.f90:226-227): one statement = one record.  The
            # values go to capture unit 9000+u in list order (a whole array: every element, first dimension fastest),
            # the record structure to unit 9500+u as  -1, (kind, count) per item  with kind = 4 (integer) or 8 (real):
            # enough to rebuild the record's bytes.  (The 4-byte record markers around them are the compiler runtime's.)
            u = int(m.group(1))
            self.emit("ref_capture(S, %d, -1.0);" % (9500 + u))
            for item in split_top(m.group(2)):
                item = item.strip()
                if not item:
                    continue
                e = parse_expr(item)
                kind = 8 if self.typeof(e) == "real" else 4
                shape = self.section_shape(e)
                if shape:
                    op, cl, names = self.loops(shape)
                    self.emit("{ long cnt_ = 0; %s{ ref_capture(S, %d, (double)(%s)); ++cnt_; }%s ref_capture(S, %d, %d.0); "
                              "ref_capture(S, %d, (double)cnt_); }" % (op, 9000 + u, self.cx(e, names), cl, 9500 + u, kind, 9500 + u))
                else:
                    self.emit("ref_capture(S, %d, (double)(%s)); ref_capture(S, %d, %d.0); ref_capture(S, %d, 1.0);"
                              % (9000 + u, self.cx(e), 9500 + u, kind, 9500 + u))
            return
        m = re.match(r"^read\s*\(\s*(\d+)\s*\)\s*(.*)$", t)
        if m:
            # UNFORMATTED sequential input (loadcntdflow, saveload.f90:327-328): the items are filled, in list order and
            # first dimension fastest, from the rank's playback queue -- which a test loads with what the matching
            # writer handed to its write(unit) (ref_world_set_playback)
            for item in split_top(m.group(2)):
                item = item.strip()
                if not item:
                    continue
                e = parse_expr(item)
                ctype = "double" if self.typeof(e) == "real" else "int"
                shape = self.section_shape(e)
                if shape:
                    op, cl, names = self.loops(shape)
                    self.emit("%s%s = (%s)ref_play(S, %s);%s" % (op, self.cx(e, names), ctype, m.group(1), cl))
                else:
                    self.emit("%s = (%s)ref_play(S, %s);" % (self.cx(e), ctype, m.group(1)))
            return
        m = re.match(r"^write\s*\(\s*(\d+)\s*,", t)
        if m:
            # formatted output to a file unit: hand the numeric items to the capture buffer in list order
            # (character items are dropped; a whole array contributes all its elements)
            close = matching_paren(t, t.index("("))
            for item in split_top(t[close + 1:]):
                item = item.strip()
                if not item or item[0] in "'\"":
                    continue
                e = parse_expr(item)
                shape = self.section_shape(e)
                if shape:
                    op, cl, names = self.loops(shape)
                    self.emit("%sref_capture(S, %s, (double)(%s));%s" % (op, m.group(1), self.cx(e, names), cl))
                else:
                    self.emit("ref_capture(S, %s, (double)(%s));" % (m.group(1), self.cx(e)))
            return
        if t.startswith("write") or t.startswith("print") or t.startswith("open") or t.startswith("close") \
                or t.startswith("format") or t.startswith("read"):
            self.emit("/* i/o statement skipped */;")
            return
        m = re.match(r"^where\s*\(", t)
        if m:
            # single-statement WHERE: masked elementwise assignment
            k = t.index("(")
            e = matching_paren(t, k)
            mask = parse_expr(t[k + 1:e])
            rest = t[e + 1:].strip()
            eq = rest.index("=")
            lhs, rhs = parse_expr(rest[:eq].strip()), parse_expr(rest[eq + 1:].strip())
            op, cl, names = self.loops(self.section_shape(lhs))
            self.emit("%sif (%s) %s = %s;%s" % (op, self.cx(mask, names), self.cx(lhs, names), self.cx(rhs, names), cl))
            return
        if t == "return":
            self.emit("goto ref_end;")
            return
        if t == "stop" or t.startswith("stop "):
            self.emit("ref_stop(S);")
            return
        m = re.match(r"^go\s*to\s+(\d+)$", t)
        if m:
            self.emit("goto L%s;" % m.group(1))
            return
        if re.match(r"^end\s*do$", t):
            self.emit("} }")
            return
        if re.match(r"^end\s*if$", t):
            self.emit("}")
            return
        if re.match(r"^end\s*select$", t):
            self.emit("break; } }" if self.in_case else "}")
            self.in_case = False
            return
        if t == "else":
            self.emit("} else {")
            return
        m = re.match(r"^else\s*if\s*\(", t)
        if m:
            k = t.index("(")
            e = matching_paren(t, k)
            self.emit("} else if (%s) {" % self.cx(parse_expr(t[k + 1:e])))
            return
        m = re.match(r"^select\s+case\s*\(", t)
        if m:
            k = t.index("(")
            e = matching_paren(t, k)
            self.emit("switch ((int)(%s)) {" % self.cx(parse_expr(t[k + 1:e])))
            self.in_case = False
            return
        m = re.match(r"^case\s*\((.*)\)$", t)
        if m:
            if self.in_case:
                self.emit("break; }")
            self.emit("case %s: {" % self.cx(parse_expr(m.group(1))))
            self.in_case = True
            return
        if t == "do":
            self.emit("{ for (;;) {")
            return
        if t == "exit":
            self.emit("break;")
            return
        m = re.match(r"^do\s+([a-z_0-9]+)\s*=\s*(.*)$", t)
        if m:
            var = self.cx(parse_expr(m.group(1)))
            parts = split_top(m.group(2))
            lo, hi = self.cx(parse_expr(parts[0])), self.cx(parse_expr(parts[1]))
            self.tmp_id += 1
            hv = "do_hi%d" % self.tmp_id
            if len(parts) == 3:
                st = self.cx(parse_expr(parts[2]))
                self.emit("{ const int %s = %s; const int %s_st = %s; for (%s = %s; (%s_st > 0) ? (%s <= %s) : (%s >= %s); %s += %s_st) {"
                          % (hv, hi, hv, st, var, lo, hv, var, hv, var, hv, var, hv))
            else:
                self.emit("{ const int %s = %s; for (%s = %s; %s <= %s; ++%s) {" % (hv, hi, var, lo, var, hv, var))
            return
        m = re.match(r"^if\s*\(", t)
        if m:
            k = t.index("(")
            e = matching_paren(t, k)
            cond = self.cx(parse_expr(t[k + 1:e]))
            rest = t[e + 1:].strip()
            if rest == "then":
                self.emit("if (%s) {" % cond)
            else:
                self.emit("if (%s) {" % cond)
                self.statement(rest, ov)
                self.emit("}")
            return
        if t.startswith("call "):
            self.call(t)
            return
        m = re.match(r"^allocate\s*\((.*)\)$", t)
        if m:
            for ent in split_top(m.group(1)):
                k = ent.index("(")
                name, dims = ent[:k].strip(), parse_dims(ent[k + 1: matching_paren(ent, k)])
                s = self.sym(name)
                if not s or s.typ not in ("real", "int"):
                    self.emit("/* allocate(%s) skipped: not a real/integer array */;" % name)
                    continue
                los = ", ".join("(%s)" % self.cx(parse_expr(lo)) for lo, hi in dims)
                his = ", ".join("(%s)" % self.cx(parse_expr(hi)) for lo, hi in dims)
                kind = "REF_KIND_R" if s.typ == "real" else "REF_KIND_I"
                self.emit("ref_free(&%s); %s = ref_alloc(%s, %d, (int[]){%s}, (int[]){%s});"
                          % (self.c_name(s), self.c_name(s), kind, len(dims), los, his))
            return
        m = re.match(r"^deallocate\s*\((.*)\)$", t)
        if m:
            for name in split_top(m.group(1)):
                sname = self.sym(name.strip())
                if sname and sname.typ in ("real", "int"):
                    self.emit("ref_free(&%s);" % self.c_name(sname))
            return
        # assignment: split at the top-level '=' that is not part of ==, /=, <=, >=
        depth, q = 0, None
        for i, ch in enumerate(t):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if t[i + 1:i + 2] == "=" or t[i - 1] in "/<>=":
                    continue
                lname = re.match(r"[a-z_0-9]+", t[:i].strip())
                ls = self.sym(lname.group(0)) if lname else None
                if ls is not None and ls.typ == "other":
                    # character assignment (file names, saveload.f90:214-218): the pieces of the concatenation go to
                    # capture unit 9900 -- a literal as its character codes, char(expr) as the value of expr, anything
                    # else (trim(directory)) as -1
                    self.emit("ref_capture(S, 9900, -2.0);")
                    for piece in split_concat(t[i + 1:].strip()):
                        piece = piece.strip()
                        if piece[:1] in "'\"":
                            for ch in piece[1:-1]:
                                self.emit("ref_capture(S, 9900, %d.0);" % ord(ch))
                        elif piece.startswith("char(") and piece.endswith(")"):
                            self.emit("ref_capture(S, 9900, (double)(%s));" % self.cx(parse_expr(piece[5:-1])))
                        else:
                            self.emit("ref_capture(S, 9900, -1.0);")
                    return
                self.assignment(t[:i].strip(), t[i + 1:].strip(), override_ok=ov)
                return
        raise SyntaxError("statement not understood: %r" % t)

    # ---- whole files -----------------------------------------------------------------------------
    def subroutines_of(self, fname):
        """{name: [logical lines]}"""
        lines = logical_lines(os.path.join(self.ref, fname))
        subs, cur, name = {}, None, None
        for no, t in lines:
            m = re.match(r"^subroutine\s+([a-z_0-9]+)", t)
            if m:
                name, cur = m.group(1), [(no, t)]
                continue
            if cur is not None:
                cur.append((no, t))
                if re.match(r"^end\s*subroutine", t):
                    subs[name] = cur
                    cur = None
        return subs

    def program_of(self, fname):
        """the PROGRAM unit of a file as the logical lines of an argument-less subroutine `main`"""
        cur = None
        for no, t in logical_lines(os.path.join(self.ref, fname)):
            if re.match(r"^program\s+[a-z_0-9]+$", t):
                cur = [(no, "subroutine main")]
            elif cur is not None and re.match(r"^end\s*program", t):
                cur.append((no, "end subroutine"))
                return cur
            elif cur is not None:
                cur.append((no, t))
        raise SyntaxError("no PROGRAM unit in %s" % fname)

    def generate(self):
        self.in_case = False
        self.wanted = ["para", "allocarray", "initpop", "initvel", "collisionexchnge", "collision_mrt", "macrovar",
                       "rhoupdat", "avedensity", "forcing", "forcingp", "exchng8", "vortcalc", "sijstat00",
                       "savecntdflow", "saveinitflow", "saveprerelax",
                       "statistc", "statistc2", "diag", "outputflow", "outputuy", "outputpress", "probe",
                       "loadcntdflow", "loadinitflow"]
        o = self.emit
        o("/* GENERATED by oracle/f90toc.py from the reference's Fortran sources -- do not edit, do not commit. */")
        o('#include "../ref_runtime.h"')
        # ---- the state struct
        o("struct ref_state {")
        o("    ref_common c;   /* must be first: the runtime sees only this part */")
        for n in self.module_order:
            s = self.module[n]
            if s.typ == "other":
                continue
            if s.dims:
                o("    ref_arr %s;" % n)
            else:
                o("    %s %s;" % ({"real": "double", "int": "int", "logical": "int"}[s.typ], n))
        o("};")
        # prototypes (collisionExchnge is called before it is defined)
        o("void ref_collisionexchnge(ref_state *S, void *a, void *b, void *c, void *d);")
        o("void ref_exchng8(ref_state *S, void *a, void *b, void *c, void *d, void *e, void *f, void *g, void *h);")
        for n in ("outputuy", "outputpress"):                   # called by outputflow before they are defined
            o("void ref_%s(ref_state *S);" % n)
        # ---- parameters and fixed-shape module arrays
        self.local, self.cur_sub = {}, "var_inc"
        o("\nvoid ref_module_init(ref_state *S)\n{")
        for n in self.module_order:
            s = self.module[n]
            if s.param is not None and s.typ in ("real", "int") and not s.dims:
                fn = "ref_override_d" if s.typ == "real" else "ref_override_i"
                o('    S->%s = %s(S, "%s", %s);' % (n, fn, n, self.cx(parse_expr(s.param))))
        for n in self.module_order:
            s = self.module[n]
            if s.dims and not s.allocatable and s.typ in ("real", "int"):
                los = ", ".join("(%s)" % self.cx(parse_expr(lo)) for lo, hi in s.dims)
                his = ", ".join("(%s)" % self.cx(parse_expr(hi)) for lo, hi in s.dims)
                kind = "REF_KIND_R" if s.typ == "real" else "REF_KIND_I"
                o("    S->%s = ref_alloc(%s, %d, (int[]){%s}, (int[]){%s});" % (n, kind, s.rank, los, his))
        o("}")
        o("\nvoid ref_module_free(ref_state *S)\n{")
        for n in self.module_order:
            s = self.module[n]
            if s.dims and s.typ in ("real", "int"):
                o("    ref_free(&S->%s);" % n)
        o("}")
        # ---- subroutines
        para = self.subroutines_of("para.f90")
        init = self.subroutines_of("initial.f90")
        coll = self.subroutines_of("collision.f90")
        # `para` is translated up to the pre-relaxation tolerance; what follows is output
        # directories (character handling) and the particle parameter block (ipart = .false.)
        self.translate_sub("para", para["para"], stop_after=r"^rhoepsl\s*=", override_assignments=True)
        self.translate_sub("allocarray", para["allocarray"])
        self.translate_sub("initpop", init["initpop"])
        self.translate_sub("initvel", init["initvel"], override_assignments=True)
        o("#ifndef REF_DROPIN")
        for n in ("collision_mrt", "collisionexchnge", "macrovar", "rhoupdat", "avedensity", "forcing", "forcingp"):
            self.translate_sub(n, coll[n])
        o("#else   /* the drop-in's collision_b200.f90 instead of the reference's collision.f90 */")
        for n in ("collision_mrt", "macrovar", "rhoupdat", "avedensity", "forcing", "forcingp"):
            o("void ref_%s(ref_state *S);" % n)
        o('#include "shim_translated.c"')
        o("#endif")
        # next-tier diagnostics: vorticity and its ghost-plane exchange
        save = self.subroutines_of("saveload.f90")
        self.translate_sub("vortcalc", save["vortcalc"])
        self.translate_sub("exchng8", save["exchng8"])
        # rank 1: profile statistics and the diag monitor; their write(unit, ...) lists are captured
        for n in ("statistc", "statistc2", "diag"):
            self.translate_sub(n, save[n])
        # what main.f90 calls every nflowout steps and after the loop: rank 0 collects uy / rho from all ranks
        # (MPI_SEND/MPI_RECV) and writes a profile and a cut; the centre-node probe of every rank (MPI_GATHER)
        for n in ("outputflow", "outputuy", "outputpress", "probe"):
            self.translate_sub(n, save[n])
        # rank 4, second half: the local strain rate from the non-equilibrium moments (sijstat00's first loop nest,
        # saveload.f90:2031-2091; what follows in that routine are particle-centred shell statistics).  Sij2 is an
        # automatic array there: its elements are captured under unit 99.
        self.translate_sub("sijstat00", save["sijstat00"], stop_after=(r"^end\s*do$", 3), capture_local=("sij2", 99))
        # rank 2: the checkpoint writers -- file name pieces and the records of their unformatted writes are captured
        for n in ("savecntdflow", "saveinitflow", "saveprerelax"):
            self.translate_sub(n, save[n])
        # ... and the reader of a continued run (main.f90:120): its read(unit) lists are served from a playback queue
        self.translate_sub("loadcntdflow", save["loadcntdflow"])
        self.translate_sub("loadinitflow", save["loadinitflow"])     # main.f90:112: a new run from a saved pre-relaxed flow
        # the driver itself: PROGRAM main (main.f90:19-236)
        self.wanted.append("main")
        self.translate_sub("main", self.program_of("main.f90"))
        # ---- dispatch + reflection tables for the Python wrapper
        o("\nint ref_dispatch(ref_state *S, const char *name)\n{")
        for n in self.wanted:
            if n in ("collisionexchnge", "exchng8"):
                continue
            o('    if (!strcmp(name, "%s")) { ref_%s(S); return 0; }' % (n, n))
        o('    if (!strcmp(name, "module_init")) { ref_module_init(S); return 0; }')
        o("    return 1;\n}")
        o("\nvoid *ref_scalar(ref_state *S, const char *name, int *kind)\n{")
        for n in self.module_order:
            s = self.module[n]
            if s.dims or s.typ == "other":
                continue
            o('    if (!strcmp(name, "%s")) { *kind = %s; return &S->%s; }'
              % (n, "REF_KIND_R" if s.typ == "real" else "REF_KIND_I", n))
        o("    return 0;\n}")
        o("\nref_arr *ref_array(ref_state *S, const char *name)\n{")
        for n in self.module_order:
            s = self.module[n]
            if s.dims and s.typ in ("real", "int"):
                o('    if (!strcmp(name, "%s")) return &S->%s;' % (n, n))
        o("    return 0;\n}")
        o("size_t ref_state_size(void) { return sizeof(ref_state); }")
        return "\n".join(self.out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/Channel-Flow")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    tr = Translator(a.ref)
    text = tr.generate()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as fh:
        fh.write(text)
    print("f90toc: wrote %s (%d lines) from %s" % (a.out, text.count("\n"), a.ref))


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""shim2c.py -- reads THIS repository's Fortran shim (d3q19-single-phase_b200/fortran/collision_b200.f90) and
(1) checks its ISO_C_BINDING declarations against include/d3q19_b200.h, (2) translates it to C.

TEST INFRASTRUCTURE ONLY (like oracle/f90toc.py, whose reader and expression parser it uses).  The image has no Fortran
compiler, so the shim a maintainer of the reference would link instead of collision.f90 cannot be compiled here.  What can
be done without one:

  * lint(): every `bind(c)` interface of the shim is compared with the prototype of the same name in the header -- number
    of arguments, by-value / by-reference passing, integer / real / pointer class and width of each, the result type --
    and the `type, bind(c) :: d3q19_config` mirror is compared with the header's struct field by field (order, type,
    array length).  An interface that does not match the C-ABI is exactly the error a Fortran link would NOT catch.
    Runs anywhere (tests/test_fortran_shim.py).
  * translate(): the executable part of the shim -- module variables, d3q19_b200_ensure, d3q19_b200_check, the seven
    argument-less subroutines main.f90 calls -- becomes C, statement by statement, against the state struct f90toc.py
    generates from the reference's var_inc.f90.  Built into oracle/_ref/libref_b200.so together with the translated
    reference MINUS its collision.f90 (f90toc.py, -DREF_DROPIN): the reference's own PROGRAM main then drives
    libd3q19b200.so through the shim, as the Fortran link line of INTEGRATION.md would have it, and is compared with the
    all-reference build (tests/test_reference_driver.py).  Needs /root/reference at build time only.

Module variables of the shim (`handle`, `bound`) exist once per MPI rank = process in the real job; here every rank is a
thread, so they live in a table indexed by rank.  A `type(d3q19_config)` local starts out filled with 0xA5 bytes, not
zeros: a field the shim forgets to set must not pass by accident.  Integer assignments to `cfg%<field>` can be overridden
by name ("cfg%math") through the translated reference's override table -- how the tests ask for STRICT arithmetic.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import f90toc as F  # noqa: E402

SHIM = os.path.join(ROOT, "d3q19-single-phase_b200", "fortran", "collision_b200.f90")
HEADER = os.path.join(ROOT, "include", "d3q19_b200.h")

KIND_C = {"c_int32_t": "int32_t", "c_int": "int", "c_double": "double", "c_signed_char": "signed char",
          "c_size_t": "size_t", "c_int64_t": "int64_t", "c_char": "char"}
# class and width a C parameter / field type belongs to: what has to agree across the binding
CLASS_OF = {"int32_t": ("int", 4), "int": ("int", 4), "int64_t": ("int", 8), "double": ("real", 8), "size_t": ("int", 8),
            "signed char": ("int", 1), "unsigned char": ("int", 1), "char": ("int", 1), "uint8_t": ("int", 1),
            "int8_t": ("int", 1), "void": ("void", 0)}


def split_semis(text):
    out, cur, q = [], "", None
    for ch in text:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch; cur += ch
        elif ch == ";":
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def statements(path):
    res = []
    for no, t in F.logical_lines(path):
        for s in split_semis(t):
            res.append((no, s))
    return res


# ------------------------------------------------------------------------------------------------------------------
# reading the shim
# ------------------------------------------------------------------------------------------------------------------
class Decl:
    def __init__(self, name, base, kind, attrs, dims):
        self.name, self.base, self.kind, self.attrs, self.dims = name, base, kind, attrs, dims
    # base: integer | real | logical | type | character ; kind: c_int32_t, c_ptr, d3q19_config, None ...

    @property
    def ctype(self):
        if self.base == "type":
            return "void *" if self.kind == "c_ptr" else self.kind
        if self.base == "character":
            return "char"
        if self.base == "logical":
            return "int"
        if self.kind is None:
            return "int" if self.base == "integer" else "double"
        return KIND_C[self.kind]


DECL_RE = re.compile(r"^(integer|real|logical|type|character)\s*(\(([^)]*)\))?\s*(.*)$")


def parse_decls(text):
    """one declaration statement -> [Decl]"""
    m = DECL_RE.match(text)
    if not m:
        return None
    base, kind, rest = m.group(1), m.group(3), m.group(4)
    if base == "type" and kind is None:
        return None
    if kind:
        kind = kind.strip()
        kind = re.sub(r"^(kind|len)\s*=\s*", "", kind)
        if base == "character":
            kind = None
    attrs = {}
    if "::" in rest:
        a, names = rest.split("::", 1)
        for item in F.split_top(a.strip().lstrip(",")):
            item = item.strip()
            if item:
                attrs[item.split("(")[0].strip()] = item
    else:
        names = rest
    out = []
    for ent in F.split_top(names):
        ent = ent.strip()
        init = None
        if "=" in ent:
            ent, init = [x.strip() for x in ent.split("=", 1)]
        dims = None
        if "(" in ent:
            k = ent.index("(")
            dims = ent[k + 1: F.matching_paren(ent, k)].strip()
            ent = ent[:k].strip()
        d = Decl(ent, base, kind, attrs, dims)
        d.init = init
        out.append(d)
    return out


class Shim:
    def __init__(self, path=SHIM):
        self.path = path
        self.params = {}          # name -> initialiser text
        self.cfg = []             # [Decl] fields of type d3q19_config, in order
        self.modvars = {}         # name -> Decl
        self.iface = {}           # fortran name -> dict(cname, ret (Decl or None), args [Decl])
        self.subs = {}            # name -> dict(dummies, decls {name: Decl}, body [(no, text)], contained)
        self._read()

    def _read(self):
        st = statements(self.path)
        i, n = 0, len(st)
        in_module, contained = False, False
        while i < n:
            no, t = st[i]
            if re.match(r"^module\s+[a-z_0-9]+$", t):
                in_module = True
            elif re.match(r"^end\s*module", t):
                in_module, contained = False, False
            elif t == "contains":
                contained = True
            elif re.match(r"^type\s*,\s*bind\s*\(\s*c\s*\)\s*::\s*d3q19_config$", t):
                i += 1
                while not re.match(r"^end\s*type", st[i][1]):
                    self.cfg.extend(parse_decls(st[i][1]))
                    i += 1
            elif t == "interface":
                i += 1
                while st[i][1] != "end interface":
                    i = self._read_interface(st, i)
            elif re.match(r"^subroutine\s", t):
                i = self._read_sub(st, i, contained)
                continue
            elif in_module and not contained:
                ds = parse_decls(t)
                if ds:
                    for d in ds:
                        if "parameter" in d.attrs:
                            self.params[d.name] = d.init
                        else:
                            self.modvars[d.name] = d
            i += 1

    def _read_interface(self, st, i):
        head = st[i][1]
        m = re.match(r"^(?:(integer|real)\s*\(([a-z_0-9]+)\)\s*)?function\s+([a-z_0-9]+)\s*\(([^)]*)\)\s*(.*)$", head)
        if not m:
            raise SyntaxError("interface header not understood: %r (line %d)" % (head, st[i][0]))
        fname, dummies, tail = m.group(3), [a.strip() for a in m.group(4).split(",") if a.strip()], m.group(5)
        b = re.search(r"bind\s*\(\s*c\s*,\s*name\s*=\s*'([^']+)'\s*\)", tail)
        if not b:
            raise SyntaxError("interface %s has no bind(c, name=...)" % fname)
        res = re.search(r"result\s*\(\s*([a-z_0-9]+)\s*\)", tail)
        ret = Decl(fname, m.group(1), m.group(2), {}, None) if m.group(1) else None
        decls = {}
        i += 1
        while not re.match(r"^end\s*function", st[i][1]):
            t = st[i][1]
            if not t.startswith("import"):
                for d in parse_decls(t) or []:
                    decls[d.name] = d
            i += 1
        if res:
            ret = decls.pop(res.group(1))
        missing = [a for a in dummies if a not in decls]
        if missing:
            raise SyntaxError("interface %s: dummy arguments without a declaration: %s" % (fname, missing))
        self.iface[fname] = dict(cname=b.group(1), ret=ret, args=[decls[a] for a in dummies])
        return i + 1

    def _read_sub(self, st, i, contained):
        m = re.match(r"^subroutine\s+([a-z_0-9]+)\s*(\(([^)]*)\))?$", st[i][1])
        name = m.group(1)
        dummies = [a.strip() for a in m.group(3).split(",")] if m.group(3) else []
        decls, body = {}, []
        i += 1
        while not re.match(r"^end\s*subroutine", st[i][1]):
            no, t = st[i]
            if t.startswith("use ") or t.startswith("implicit "):
                pass
            elif not body and parse_decls(t):
                for d in parse_decls(t):
                    decls[d.name] = d
            else:
                body.append((no, t))
            i += 1
        self.subs[name] = dict(dummies=dummies, decls=decls, body=body, contained=contained)
        return i + 1


# ------------------------------------------------------------------------------------------------------------------
# the header
# ------------------------------------------------------------------------------------------------------------------
def read_header(path=HEADER):
    """-> (prototypes {name: (ret, [(type text, name)])}, fields of d3q19_config [(type, name, array length or None)])"""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9 \*]*?)\b(d3q19_[a-z_0-9]+)\s*\(([^()]*)\)\s*;", text):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), m.group(3)
        al = []
        for a in [x.strip() for x in args.split(",")]:
            if a in ("", "void"):
                continue
            am = re.match(r"^(.*?)([A-Za-z_][A-Za-z_0-9]*)\s*(\[[^\]]*\])?$", a)
            typ = " ".join(am.group(1).split())
            if am.group(3):
                typ += " *"
            al.append((typ, am.group(2)))
        protos[name] = (ret, al)
    sm = re.search(r"typedef\s+struct\s+d3q19_config\s*\{(.*?)\}\s*d3q19_config\s*;", text, flags=re.S)
    fields = []
    for decl in sm.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        tm = re.match(r"^((?:unsigned |signed )?[A-Za-z_0-9]+)\s+(.*)$", decl)
        for ent in tm.group(2).split(","):
            ent = ent.strip()
            am = re.match(r"^([A-Za-z_0-9]+)\s*(?:\[(\d+)\])?$", ent)
            fields.append((tm.group(1), am.group(1), int(am.group(2)) if am.group(2) else None))
    return protos, fields


def c_param_class(typ):
    """C parameter type -> (class, width, pointer depth)"""
    t = typ.replace("const", " ")
    depth = t.count("*")
    base = " ".join(t.replace("*", " ").split())
    if base in CLASS_OF:
        return CLASS_OF[base] + (depth,)
    return ("struct:" + base, 0, depth)


def f_param_class(d):
    """Fortran dummy -> (class, width, pointer depth) as the companion C processor sees it"""
    by_value = "value" in d.attrs
    if d.base == "type":
        base = ("void", 0, 1) if d.kind == "c_ptr" else ("struct:" + d.kind, 0, 0)
    else:
        cls, w = CLASS_OF[d.ctype]
        base = (cls, w, 0)
    if d.dims is not None or not by_value:
        return (base[0], base[1], base[2] + 1)
    return base


def lint(shim=None, header=HEADER):
    """-> list of human-readable mismatches between the shim's ISO_C_BINDING declarations and the header (empty = fine)"""
    shim = shim or Shim()
    protos, fields = read_header(header)
    bad = []
    # ---- the struct mirror
    if len(shim.cfg) != len(fields):
        bad.append("d3q19_config: %d fields in the shim, %d in the header" % (len(shim.cfg), len(fields)))
    for d, (ctyp, cname, clen) in zip(shim.cfg, fields):
        flen = int(d.dims) if d.dims else None
        if d.name != cname.lower():
            bad.append("d3q19_config: field %r in the shim where the header has %r" % (d.name, cname))
        elif CLASS_OF[d.ctype] != CLASS_OF[ctyp] or flen != clen:
            bad.append("d3q19_config.%s: %s%s in the shim, %s%s in the header"
                       % (cname, d.ctype, "[%s]" % flen if flen else "", ctyp, "[%s]" % clen if clen else ""))
    # ---- every interface against the prototype of its bind name
    for fname, it in sorted(shim.iface.items()):
        cname = it["cname"]
        if cname == "strlen":                   # libc, bound under another Fortran name on purpose
            continue
        if fname != cname.lower():
            bad.append("%s: bind name %r differs from the Fortran name" % (fname, cname))
        if cname not in protos:
            bad.append("%s: no such entry point in the header" % cname)
            continue
        ret, cargs = protos[cname]
        fret = f_param_class(Decl("", it["ret"].base, it["ret"].kind, {"value": "value"}, None)) if it["ret"] else None
        cret = c_param_class(ret)
        if fret is None or (fret[0] != cret[0] and not (fret[0] == "void" and cret[2] == 1)) or fret[2] != cret[2] \
                or (fret[2] == 0 and fret[1] != cret[1]):
            bad.append("%s: result %s in the shim, `%s` in the header" % (cname, fret, ret))
        if len(cargs) != len(it["args"]):
            bad.append("%s: %d arguments in the shim, %d in the header" % (cname, len(it["args"]), len(cargs)))
            continue
        for d, (ctyp, an) in zip(it["args"], cargs):
            fc, cc = f_param_class(d), c_param_class(ctyp)
            ok = fc[2] == cc[2]
            if ok and fc[2] == 0:
                ok = fc[0] == cc[0] and fc[1] == cc[1]                    # by value: same class and width
            elif ok and d.base == "type" and d.kind == "c_ptr":
                ok = True                                                 # an opaque pointer (or the address of one)
            elif ok:
                # by reference: same pointee class and width; a byte buffer may be `void *` on the C side
                ok = (fc[0] == cc[0] and fc[1] == cc[1]) or (cc[0] == "void" and fc[1] == 1)
            if not ok:
                bad.append("%s: argument %r is %s in the shim, `%s %s` in the header" % (cname, d.name, fc, ctyp, an))
    # ---- the constants the shim mirrors
    text = re.sub(r"/\*.*?\*/", " ", open(header).read(), flags=re.S)
    for name, val in shim.params.items():
        m = re.search(r"\b%s\b\s*=?\s*\(?(-?\d+)" % name.upper(), text)
        if not m:
            bad.append("parameter %s: not found in the header" % name.upper())
        elif int(m.group(1)) != int(val):
            bad.append("parameter %s = %s in the shim, %s in the header" % (name.upper(), val, m.group(1)))
    return bad


# ------------------------------------------------------------------------------------------------------------------
# translation
# ------------------------------------------------------------------------------------------------------------------
class ShimTranslator:
    def __init__(self, ref_dir, shim=None):
        self.shim = shim or Shim()
        self.tr = F.Translator(ref_dir)
        if not self.tr.module:
            self.tr.parse_module()
        self.protos, self.fields = read_header()
        self.out = []
        self.local = {}
        self.cur = ""

    def emit(self, s):
        self.out.append(s)

    # ---- names ---------------------------------------------------------------------------------------------------
    @staticmethod
    def prep(text):
        """`cfg%lx` -> one identifier, kind suffixes of literals dropped"""
        text = re.sub(r"\b([a-z_][a-z0-9_]*)%([a-z_][a-z0-9_]*)", r"\1__\2", text)
        return re.sub(r"\b(\d+)_c_[a-z0-9_]+\b", r"\1", text)

    def cfg_field(self, name):
        if "__" in name:
            var, fld = name.split("__", 1)
            d = self.local.get(var)
            if d and d.base == "type" and d.kind == "d3q19_config":
                for f in self.shim.cfg:
                    if f.name == fld:
                        return var, f
                raise KeyError("d3q19_config has no field %r (%s)" % (fld, self.cur))
        return None

    def var(self, name):
        """-> (C text, kind) with kind in scalar | array | struct | charptr"""
        cf = self.cfg_field(name)
        if cf:
            return "%s.%s" % (cf[0], cf[1].name), ("array" if cf[1].dims else "scalar")
        if name in self.local:
            d = self.local[name]
            if d.base == "type" and d.kind != "c_ptr":
                return name, "struct"
            if d.base == "character":
                return name, "charptr"
            return name, ("array" if d.dims else "scalar")
        if name in self.shim.modvars:
            return "M->%s" % name, "scalar"
        if name in self.shim.params:
            return "(%s)" % self.shim.params[name], "const"
        s = self.tr.module.get(name)
        if s is not None:
            if s.dims:
                return "S->%s.p" % name, "array"
            return "S->%s" % name, "scalar"
        raise KeyError("unknown name %r in %s" % (name, self.cur))

    # ---- expressions -----------------------------------------------------------------------------------------------
    def cx(self, e):
        if isinstance(e, F.Num):
            return self.tr.num_c(e)
        if isinstance(e, F.Logical):
            return str(e.v)
        if isinstance(e, F.Var):
            if e.name in F.MPI_CONST:
                return F.MPI_CONST[e.name]
            if e.name == "c_null_ptr":
                return "((void *)0)"
            if e.name in self.shim.iface:                        # a function without arguments
                return self.c_call(e.name, [])
            return self.var(e.name)[0]
        if isinstance(e, F.Un):
            if e.op == "()":
                return "(%s)" % self.cx(e.e)
            return "(%s(%s))" % (e.op, self.cx(e.e))
        if isinstance(e, F.Bin):
            return "(%s %s %s)" % (self.cx(e.l), e.op, self.cx(e.r))
        if isinstance(e, F.Index):
            n = e.name
            if n in self.shim.iface:
                return self.c_call(n, e.args)
            a = e.args
            if n == "merge":
                return "((%s) ? (%s) : (%s))" % (self.cx(a[2]), self.cx(a[0]), self.cx(a[1]))
            if n == "mod":
                return "((%s) %% (%s))" % (self.cx(a[0]), self.cx(a[1]))
            if n in ("max", "min"):
                x, y = self.cx(a[0]), self.cx(a[1])
                return "((%s) %s (%s) ? (%s) : (%s))" % (x, ">" if n == "max" else "<", y, x, y)
            if n == "int":
                return "((%s)(%s))" % (KIND_C.get(a[1].name, "int") if len(a) > 1 else "int", self.cx(a[0]))
            if n == "real":
                return "((double)(%s))" % self.cx(a[0])
            raise KeyError("unsupported function %r in %s" % (n, self.cur))
        raise TypeError("cannot translate %r in %s" % (e, self.cur))

    def c_call(self, fname, actuals):
        it = self.shim.iface[fname]
        cname = it["cname"]
        if len(actuals) != len(it["args"]):
            raise SyntaxError("%s called with %d arguments, interface has %d (%s)" % (fname, len(actuals), len(it["args"]), self.cur))
        if cname == "strlen":
            return "strlen((const char *)%s)" % self.cx(actuals[0])
        htypes = [t for t, _ in self.protos[cname][1]]
        parts = []
        for d, a, ht in zip(it["args"], actuals, htypes):
            if "value" in d.attrs and d.dims is None:
                parts.append("(%s)(%s)" % (ht, self.cx(a)))
                continue
            # by reference
            if isinstance(a, F.Var) and a.name not in F.MPI_CONST:
                text, kind = self.var(a.name)
                if kind in ("array", "charptr"):
                    parts.append("(%s)%s" % (ht, text))
                    continue
                if kind in ("scalar", "struct"):
                    parts.append("(%s)&%s" % (ht, text))
                    continue
            parts.append("(%s)&(%s){%s}" % (ht, d.ctype, self.cx(a)))
        ret = it["ret"]
        cast = "(void *)" if ret is not None and ret.base == "type" and ret.kind == "c_ptr" else ""
        return "%s%s(%s)" % (cast, cname, ", ".join(parts))

    def mpi_arg(self, a):
        e = F.parse_expr(self.prep(a))
        if isinstance(e, F.Var):
            if e.name in F.MPI_CONST:
                return F.MPI_CONST[e.name]
            text, kind = self.var(e.name)
            return text if kind == "array" else "&%s" % text
        return "&(int){%s}" % self.cx(e)

    # ---- statements ------------------------------------------------------------------------------------------------
    def statement(self, t):
        if t == "return":
            self.emit("return;")
            return
        if t in ("endif", "end if"):
            self.emit("}")
            return
        if t == "else":
            self.emit("} else {")
            return
        m = re.match(r"^if\s*\(", t)
        if m:
            k = t.index("(")
            e = F.matching_paren(t, k)
            cond = self.cx(F.parse_expr(self.prep(t[k + 1:e])))
            rest = t[e + 1:].strip()
            self.emit("if (%s) {" % cond)
            if rest != "then":
                self.statement(rest)
                self.emit("}")
            return
        m = re.match(r"^write\s*\(\s*\*\s*,\s*\*\s*\)\s*(.*)$", t)
        if m:
            fmt, args = "", []
            for item in F.split_top(m.group(1)):
                item = item.strip()
                if item[:1] in "'\"":
                    fmt += " %s"; args.append('"%s"' % item[1:-1].replace("\\", "\\\\").replace('"', '\\"'))
                    continue
                sm = re.match(r"^([a-z_0-9]+)\s*\(\s*1\s*:\s*([a-z_0-9]+)\s*\)$", item)
                if sm:
                    fmt += " %.*s"; args += ["(int)%s" % self.var(sm.group(2))[0], self.var(sm.group(1))[0]]
                    continue
                text, kind = self.var(item)
                if kind != "charptr":
                    raise SyntaxError("write item %r is not character (%s)" % (item, self.cur))
                fmt += " %s"; args.append(text)
            self.emit('fprintf(stderr, "%s\\n", %s);' % (fmt, ", ".join(args)))
            return
        m = re.match(r"^allocate\s*\(\s*([a-z_0-9]+)\s*\((.*)\)\s*\)$", t)
        if m:
            d = self.local[m.group(1)]
            self.emit("%s = (%s *)malloc((size_t)(%s) * sizeof(%s));"
                      % (d.name, d.ctype, self.cx(F.parse_expr(self.prep(m.group(2)))), d.ctype))
            return
        m = re.match(r"^deallocate\s*\(\s*([a-z_0-9]+)\s*\)$", t)
        if m:
            self.emit("free(%s); %s = 0;" % (m.group(1), m.group(1)))
            return
        m = re.match(r"^call\s+([a-z_0-9]+)\s*(\((.*)\))?$", t)
        if m:
            name, args = m.group(1), F.split_top(m.group(3)) if m.group(3) else []
            if name.startswith("mpi_"):
                self.emit("ref_%s(S, %s);" % (name, ", ".join(self.mpi_arg(a) for a in args)))
                return
            if name == "c_f_pointer":
                self.emit("%s = (const char *)%s;" % (self.var(args[1].strip())[0], self.var(args[0].strip())[0]))
                return
            if name in self.shim.subs:
                sub = self.shim.subs[name]
                if len(args) != len(sub["dummies"]):
                    raise SyntaxError("call %s: %d arguments for %d dummies" % (name, len(args), len(sub["dummies"])))
                cargs = ["S"]
                for a in args:
                    a = a.strip()
                    if a[:1] in "'\"":
                        cargs.append('"%s"' % a[1:-1])
                    else:
                        cargs.append(self.cx(F.parse_expr(self.prep(a))))
                self.emit("%s(%s);" % (self.c_sub_name(name), ", ".join(cargs)))
                return
            raise KeyError("call to %r: neither MPI, nor a subroutine of the shim (%s)" % (name, self.cur))
        # assignment
        depth = 0
        for i, ch in enumerate(t):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and t[i + 1:i + 2] != "=" and t[i - 1] not in "/<>=":
                lhs, rhs = self.prep(t[:i].strip()), self.prep(t[i + 1:].strip())
                text, kind = self.var(lhs)
                r = self.cx(F.parse_expr(rhs))
                cf = self.cfg_field(lhs)
                if kind == "array":
                    nlen = cf[1].dims if cf else self.local[lhs].dims
                    self.emit("for (int i_ = 0; i_ < (%s); ++i_) %s[i_] = %s;" % (nlen, text, r))
                elif cf and cf[1].base == "integer":
                    self.emit('%s = ref_override_i(S, "%s", %s);' % (text, lhs.replace("__", "%"), r))
                else:
                    self.emit("%s = %s;" % (text, r))
                return
        raise SyntaxError("statement not understood: %r (%s)" % (t, self.cur))

    def c_sub_name(self, name):
        return ("shim_%s" if self.shim.subs[name]["contained"] else "ref_%s") % name

    def signature(self, name):
        sub = self.shim.subs[name]
        parts = ["ref_state *S"]
        for dn in sub["dummies"]:
            d = sub["decls"][dn]
            parts.append("const char *%s" % dn if d.base == "character" else "%s %s" % (d.ctype, dn))
        return "%svoid %s(%s)" % ("static " if sub["contained"] else "", self.c_sub_name(name), ", ".join(parts))

    def translate_sub(self, name):
        sub = self.shim.subs[name]
        self.cur, self.local = name, dict(sub["decls"])
        self.emit("\n%s\n{" % self.signature(name))
        self.emit("    shim_module *M = &shim_mod[((ref_common *)S)->rank]; (void)M;")
        for dn, d in sub["decls"].items():
            if dn in sub["dummies"]:
                continue
            if d.base == "type" and d.kind == "d3q19_config":
                self.emit("d3q19_config %s; memset(&%s, 0xA5, sizeof %s);   /* undefined until assigned, as in Fortran */" % (dn, dn, dn))
            elif d.base == "character":
                self.emit("const char *%s = 0; (void)%s;" % (dn, dn))
            elif "allocatable" in d.attrs:
                self.emit("%s *%s = 0; (void)%s;" % (d.ctype, dn, dn))
            elif d.dims:
                self.emit("%s %s[%s]; (void)%s;" % (d.ctype, dn, d.dims, dn))
            elif d.base == "type":
                self.emit("void *%s = 0; (void)%s;" % (dn, dn))
            else:
                self.emit("%s %s = 0; (void)%s;" % (d.ctype, dn, dn))
        for no, t in sub["body"]:
            try:
                self.statement(t)
            except Exception as ex:
                raise type(ex)("%s (collision_b200.f90 line %d: %r)" % (ex, no, t)) from ex
        self.emit("}")

    def generate(self):
        bad = lint(self.shim)
        if bad:
            raise SystemExit("shim2c: the shim does not match include/d3q19_b200.h:\n  " + "\n  ".join(bad))
        # what a Fortran compiler would reject and a translator would not notice: a local declaration of a name that is
        # use-associated (var_inc or the shim's own module), or one name exported by both modules
        vi = set(self.tr.module)
        mod = set(self.shim.modvars) | set(self.shim.params) | set(self.shim.iface) | {"d3q19_config"} \
            | {n for n, sub in self.shim.subs.items() if sub["contained"]}
        clash = ["module d3q19_b200_shim and var_inc both export %r" % n for n in sorted(mod & vi)]
        for name, sub in self.shim.subs.items():
            local = set(sub["decls"]) - set(sub["dummies"])
            clash += ["%s declares %r, which it also gets from a module" % (name, n) for n in sorted(local & (vi | mod))]
        if clash:
            raise SystemExit("shim2c: name clashes in collision_b200.f90:\n  " + "\n  ".join(clash))
        o = self.emit
        o("/* GENERATED by oracle/shim2c.py from d3q19-single-phase_b200/fortran/collision_b200.f90 -- do not edit, do not")
        o(" * commit.  Included by ref_translated.c under -DREF_DROPIN (struct ref_state is complete there). */")
        o('#include "../../include/d3q19_b200.h"')
        o("#ifndef REF_MAX_RANKS\n#define REF_MAX_RANKS 64\n#endif")
        o("/* module d3q19_b200_shim's variables: one instance per MPI rank (a process in the real job, a thread here) */")
        o("typedef struct shim_module {")
        for n, d in self.shim.modvars.items():
            o("    %s %s;" % (d.ctype, n))
        o("} shim_module;")
        o("static shim_module shim_mod[REF_MAX_RANKS];")
        o("/* for the tests: the handle the shim created on a rank (to destroy it), and a reset between runs */")
        o("void *ref_shim_handle(ref_state *S) { return shim_mod[((ref_common *)S)->rank].handle; }")
        o("void ref_shim_reset(void) { memset(shim_mod, 0, sizeof shim_mod); }")
        for name in self.shim.subs:
            o("%s;" % self.signature(name))
        for name in self.shim.subs:
            self.translate_sub(name)
        return "\n".join(self.out) + "\n"


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/Channel-Flow")
    ap.add_argument("--out")
    ap.add_argument("--lint", action="store_true")
    a = ap.parse_args()
    if a.lint or not a.out:
        bad = lint()
        print("\n".join(bad) if bad else "shim2c: collision_b200.f90 matches include/d3q19_b200.h")
        return 1 if bad else 0
    text = ShimTranslator(a.ref).generate()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as fh:
        fh.write(text)
    print("shim2c: wrote %s (%d lines)" % (a.out, text.count("\n")))
    return 0


if __name__ == "__main__":
    sys.exit(main())

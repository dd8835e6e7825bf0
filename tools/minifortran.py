"""Tiny interpreter for the Fortran subset used by the RRTMG *setup* routines of the reference.

Build-time tool only (runs in the authoring container, reads /root/reference, never shipped to the GPU box).

The reference fills its k-distribution tables by executing ~70 small Fortran routines
(ifsrrtm/rrtm_kgb*.F90, rrtm_cmbgb*.F90, srtm_kgb*.F90, srtm_cmbgb*.F90, surrtpk/surrtrf/surrtftr/susrtm.F90,
rrtm_init_140gp.F90, srtm_init.F90).  No Fortran compiler exists in this container, so instead of re-typing
tens of thousands of literals we *interpret* those files where they lie: module declarations become numpy
arrays with Fortran bounds, array-constructor assignments / DO loops / IF blocks / unformatted READs are
transpiled statement by statement to Python and exec'd.  Arithmetic is IEEE double in the same operation
order as the source, so the resulting tables are bit-identical to what a double-precision build holds in
memory (literals without a kind suffix are rounded to single first, as Fortran does).
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np


class FArray:
    """numpy array with Fortran lower bounds and inclusive slices, column-major storage."""

    def __init__(self, bounds, dtype):
        self.lb = [b[0] for b in bounds]
        shape = [b[1] - b[0] + 1 for b in bounds]
        self.a = np.zeros(shape, dtype=dtype, order="F")

    def _idx(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        out = []
        for k, lb, n in zip(key, self.lb, self.a.shape):
            if isinstance(k, slice):
                start = 0 if k.start is None else k.start - lb
                stop = n if k.stop is None else k.stop - lb + 1
                out.append(slice(start, stop))
            else:
                i = int(k) - lb
                if i < 0 or i >= n:
                    raise IndexError(f"index {k} out of bounds {lb}:{lb + n - 1}")
                out.append(i)
        return tuple(out)

    def __getitem__(self, key):
        return self.a[self._idx(key)]

    def __setitem__(self, key, val):
        idx = self._idx(key)
        tgt = self.a[idx]
        if isinstance(tgt, np.ndarray):
            v = np.asarray(val)
            if v.ndim == 1 and tgt.ndim > 1:
                v = v.reshape(tgt.shape, order="F")
            self.a[idx] = v
        else:
            self.a[idx] = val

    def assign(self, val):
        v = np.asarray(val.a if isinstance(val, FArray) else val)
        if v.ndim == 0:
            self.a[...] = v
        elif v.ndim == 1 and self.a.ndim > 1:
            self.a[...] = v.reshape(self.a.shape, order="F")
        else:
            self.a[...] = v


def _ac(*items):
    out = []
    for it in items:
        if isinstance(it, (list, tuple, np.ndarray)):
            out.extend(list(np.ravel(it)))
        else:
            out.append(it)
    return np.array(out)


def _fint(x):
    return int(x)  # Fortran INT truncates toward zero, as Python int()


def _fmod(a, b):
    return np.fmod(a, b)


def _freal(x, *kind):
    if isinstance(x, FArray):
        return x.a.astype(np.float64)
    return np.float64(x)


_KIND_DOUBLE = ("_JPRB", "_JPRD", "_jprb", "_jprd")
_LIT = re.compile(
    r"(?<![A-Za-z0-9_.])((?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][-+]?\d+)?)(_JPRB|_JPRD|_jprb|_jprd|_JPIM|_JPRM)?(?![A-Za-z0-9_.])"
)


def _conv_literals(expr):
    def rep(m):
        txt, kind = m.group(1), m.group(2)
        is_real = ("." in txt) or ("e" in txt.lower()) or ("d" in txt.lower())
        if not is_real:
            return txt
        t = txt.lower().replace("d", "e")
        if kind in _KIND_DOUBLE or "d" in txt.lower():
            return repr(float(t))
        # default-kind real literal: single precision, promoted when used
        return repr(float(np.float32(float(t))))

    return _LIT.sub(rep, expr)


def clean_source(path):
    """Return logical statements (comments stripped, continuations joined, upper-cased outside strings)."""
    stmts = []
    cur = ""
    with open(path, "r", errors="replace") as f:
        for raw in f:
            line = raw.rstrip("\n")
            if line.lstrip().startswith("#"):
                continue
            # strip comments (no '!' inside strings in these files except formats we drop anyway)
            out = []
            q = None
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
            line = "".join(out).strip()
            if not line:
                continue
            if line.startswith("&"):
                line = line[1:].lstrip()
            cont = line.endswith("&")
            if cont:
                line = line[:-1].rstrip()
            cur += (" " if cur else "") + line
            if not cont:
                stmts.append(cur)
                cur = ""
    if cur:
        stmts.append(cur)
    return stmts


def _split_top(s, sep=","):
    parts, depth, cur = [], 0, ""
    i = 0
    while i < len(s):
        ch = s[i]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
        i += 1
    if cur.strip():
        parts.append(cur.strip())
    return parts


class Interp:
    def __init__(self, srcdirs, data_dir):
        self.srcdirs = srcdirs
        self.data_dir = data_dir
        self.modules = {}  # name -> dict
        self.units = {}  # unit number -> open file
        self.skip_calls = {"DR_HOOK", "MPL_BROADCAST", "ABOR1", "MODIFY_WV_CONTINUUM"}
        self.builtin_modules = {
            "PARKIND1": {"JPIM": 4, "JPRB": 8, "JPRD": 8, "JPRM": 4},
            "YOMHOOK": {"LHOOK": False, "JPHOOK": 8},
            "YOMLUN_ECRAD": {"NULRAD": 25, "NULOUT": 6},
            "MPL_MODULE": {},
            "YOMTAG": {"MTAGRAD": 0},
            "YOMMP0_IFSAUX": {"NPROC": 1, "MYPROC": 1},
        }

    # ---------------------------------------------------------------- files
    def find(self, name):
        for d in self.srcdirs:
            p = os.path.join(d, name.lower() + ".F90")
            if os.path.exists(p):
                return p
        raise FileNotFoundError(name)

    # -------------------------------------------------------------- modules
    def module(self, name):
        name = name.upper()
        if name in self.modules:
            return self.modules[name]
        if name in self.builtin_modules:
            self.modules[name] = dict(self.builtin_modules[name])
            return self.modules[name]
        ns = {}
        self.modules[name] = ns
        stmts = clean_source(self.find(name))
        for st in stmts:
            self._decl(st, ns)
        return ns

    def _eval_const(self, expr, ns):
        return eval(_conv_literals(expr.upper()), {"__builtins__": {}}, ns)

    def _decl(self, st, ns):
        u = st.upper().strip()
        m = re.match(r"USE\s+(\w+)\s*(?:,\s*ONLY\s*:\s*(.*))?$", u)
        if m:
            mod = self.module(m.group(1))
            if m.group(2):
                for item in _split_top(m.group(2)):
                    if "=>" in item:
                        loc, rem = [x.strip() for x in item.split("=>")]
                    else:
                        loc = rem = item.strip()
                    if rem in mod:
                        ns[loc] = mod[rem]
                    else:
                        ns.setdefault("__lazy__", {})[loc] = (mod, rem)
            else:
                ns.update({k: v for k, v in mod.items() if not k.startswith("__")})
            return True
        m = re.match(r"(INTEGER|REAL|LOGICAL|CHARACTER)\s*(\([^)]*\))?\s*((?:,\s*[A-Z]+(?:\([^)]*\))?\s*)*)::\s*(.*)$", u)
        if not m:
            return False
        typ, attrs, ents = m.group(1), m.group(3), m.group(4)
        if typ == "CHARACTER":
            return True
        is_param = "PARAMETER" in attrs
        dim_attr = re.search(r"DIMENSION\s*\(([^)]*)\)", attrs)
        dtype = np.float64 if typ == "REAL" else (np.int64 if typ == "INTEGER" else np.bool_)
        for ent in _split_top(ents):
            if is_param:
                nm, val = [x.strip() for x in ent.split("=", 1)]
                ns[nm] = self._eval_const(val, ns)
                continue
            m2 = re.match(r"(\w+)\s*(?:\((.*)\))?$", ent)
            nm, dims = m2.group(1), m2.group(2)
            if dims is None and dim_attr:
                dims = dim_attr.group(1)
            if "INTENT" in attrs:
                continue
            if dims:
                bounds = []
                for d in _split_top(dims):
                    if ":" in d:
                        lo, hi = d.split(":")
                        bounds.append((int(self._eval_const(lo, ns)), int(self._eval_const(hi, ns))))
                    else:
                        bounds.append((1, int(self._eval_const(d, ns))))
                ns[nm] = FArray(bounds, dtype)
            else:
                ns[nm] = dtype(0)
        return True

    # ------------------------------------------------------------- transpile
    def _expr(self, e, ns):
        e = e.strip()
        e = _conv_literals(e)
        # array constructors
        e = e.replace("(/", "_ac(").replace("/)", ")")
        e = re.sub(r"\.AND\.", " and ", e)
        e = re.sub(r"\.OR\.", " or ", e)
        e = re.sub(r"\.NOT\.", " not ", e)
        e = re.sub(r"\.TRUE\.", " True ", e)
        e = re.sub(r"\.FALSE\.", " False ", e)
        e = e.replace("/=", "!=")
        return self._brackets(e, ns)

    def _brackets(self, e, ns):
        """Turn NAME(...) into NAME[...] for arrays; map intrinsics."""
        out = ""
        i = 0
        n = len(e)
        while i < n:
            m = re.match(r"[A-Za-z_]\w*", e[i:])
            if m and (i == 0 or not (e[i - 1].isalnum() or e[i - 1] == "_" or e[i - 1] == ".")):
                name = m.group(0)
                j = i + len(name)
                k = j
                while k < n and e[k] == " ":
                    k += 1
                if k < n and e[k] == "(":
                    # find matching paren
                    depth, p = 0, k
                    while p < n:
                        if e[p] == "(":
                            depth += 1
                        elif e[p] == ")":
                            depth -= 1
                            if depth == 0:
                                break
                        p += 1
                    inner = e[k + 1 : p]
                    obj = ns.get(name)
                    if isinstance(obj, FArray):
                        args = [self._index(a, ns) for a in _split_top(inner)]
                        out += f"{name}[{', '.join(args)}]"
                    else:
                        fn = {"INT": "_fint", "MOD": "_fmod", "REAL": "_freal", "MIN": "min", "MAX": "max",
                              "SQRT": "_sqrt", "PRESENT": "_present", "TRIM": "_trim", "_ac": "_ac"}.get(name, name)
                        if name == "PRESENT":
                            out += "False"
                        else:
                            args = [self._brackets(a, ns) for a in _split_top(inner)]
                            out += f"{fn}({', '.join(args)})"
                    i = p + 1
                    continue
                out += name
                i = j
                continue
            out += e[i]
            i += 1
        return out

    def _index(self, a, ns):
        a = a.strip()
        if a == ":":
            return "slice(None, None)"
        # top-level colon => slice
        depth = 0
        for pos, ch in enumerate(a):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == ":" and depth == 0:
                lo, hi = a[:pos].strip(), a[pos + 1 :].strip()
                lo = self._brackets(lo, ns) if lo else "None"
                hi = self._brackets(hi, ns) if hi else "None"
                return f"slice({lo}, {hi})"
        return self._brackets(a, ns)

    def run(self, subname, **extra):
        """Execute subroutine <subname> (file <subname>.F90)."""
        stmts = clean_source(self.find(subname))
        ns = {}
        body = []
        in_sub = False
        for st in stmts:
            u = st.upper().strip()
            if re.match(r"SUBROUTINE\s+" + subname.upper() + r"\b", u):
                in_sub = True
                continue
            if not in_sub:
                continue
            if re.match(r"END\s+SUBROUTINE", u):
                break
            if u.startswith("IMPLICIT") or u.startswith("ASSOCIATE") or u.startswith("END ASSOCIATE"):
                continue
            if self._decl(st, ns):
                continue
            body.append(u)
        ns.update(extra)
        code = self._transpile(body, ns)
        env = {
            "_ac": _ac, "_fint": _fint, "_fmod": _fmod, "_freal": _freal, "_sqrt": np.sqrt,
            "_present": lambda *a: False, "_trim": lambda s: s, "_call": self._call, "_read": self._read,
            "_open": self._open, "np": np, "min": min, "max": max,
        }
        # scalars are rebound, so exec in a dict that starts from ns and copy scalar results back to modules
        g = dict(env)
        g.update({k: v for k, v in ns.items() if not k.startswith("__")})
        try:
            exec(compile(code, subname, "exec"), g)
        except Exception:
            print(code)
            raise
        self._writeback(ns, g)
        return g

    def _writeback(self, ns, g):
        # scalar module variables assigned in the routine must be propagated to their modules
        for mod in self.modules.values():
            for k, v in list(mod.items()):
                if k.startswith("__") or isinstance(v, FArray):
                    continue
                if k in ns and k in g and not isinstance(g[k], FArray) and ns[k] is mod[k]:
                    mod[k] = g[k]

    def _transpile(self, body, ns):
        lines = []
        ind = 0

        def emit(s):
            lines.append("    " * ind + s)

        for u in body:
            # drop statement labels
            u = re.sub(r"^\d+\s+", "", u)
            if u in ("CONTINUE", "RETURN") or u.startswith("WRITE") or u.startswith("FORMAT") or u.startswith("CLOSE"):
                if u == "RETURN":
                    pass
                continue
            m = re.match(r"DO\s+(\w+)\s*=\s*(.*)$", u)
            if m:
                parts = _split_top(m.group(2))
                lo, hi = self._expr(parts[0], ns), self._expr(parts[1], ns)
                emit(f"for {m.group(1)} in range({lo}, ({hi})+1):")
                ind += 1
                emit("pass")
                continue
            if re.match(r"END\s*DO$", u):
                ind -= 1
                continue
            m = re.match(r"IF\s*\((.*)\)\s*THEN$", u)
            if m:
                emit(f"if {self._expr(m.group(1), ns)}:")
                ind += 1
                emit("pass")
                continue
            m = re.match(r"ELSE\s*IF\s*\((.*)\)\s*THEN$", u)
            if m:
                ind -= 1
                emit(f"elif {self._expr(m.group(1), ns)}:")
                ind += 1
                emit("pass")
                continue
            if u == "ELSE":
                ind -= 1
                emit("else:")
                ind += 1
                emit("pass")
                continue
            if re.match(r"END\s*IF$", u):
                ind -= 1
                continue
            m = re.match(r"IF\s*\(", u)
            if m:
                # one-line IF: find matching paren
                depth, p = 0, u.index("(")
                while True:
                    if u[p] == "(":
                        depth += 1
                    elif u[p] == ")":
                        depth -= 1
                        if depth == 0:
                            break
                    p += 1
                cond, rest = u[u.index("(") + 1 : p], u[p + 1 :].strip()
                emit(f"if {self._expr(cond, ns)}:")
                ind += 1
                for s in self._simple(rest, ns):
                    emit(s)
                ind -= 1
                continue
            for s in self._simple(u, ns):
                emit(s)
        return "\n".join(lines) + "\n"

    def _simple(self, u, ns):
        m = re.match(r"CALL\s+(\w+)\s*(?:\((.*)\))?$", u)
        if m:
            if m.group(1) in self.skip_calls:
                return ["pass"]
            return [f"_call('{m.group(1)}')"]
        m = re.match(r"READ\s*\(([^)]*)\)\s*(.*)$", u)
        if m:
            names = [x.strip() for x in _split_top(m.group(2))]
            return [f"_read(NULRAD, [{', '.join(names)}])"]
        if u.startswith("OPEN"):
            fname = "RADRRTM" if "RRTM" in ns.get("__subname__", "") else None
            return ["_open(NULRAD, CLF1)"]
        m = re.match(r"(\w+)\s*(\(.*?\))?\s*=\s*(?!=)(.*)$", u)
        if m and not u.startswith("CLF1"):
            # need the *top-level* '=' : re-split carefully
            lhs, rhs = self._split_assign(u)
            rhs_py = self._expr(rhs, ns)
            mm = re.match(r"(\w+)\s*(?:\((.*)\))?$", lhs.strip())
            name, args = mm.group(1), mm.group(2)
            obj = ns.get(name)
            if isinstance(obj, FArray):
                if args is None:
                    return [f"{name}.assign({rhs_py})"]
                idx = ", ".join(self._index(a, ns) for a in _split_top(args))
                return [f"{name}[{idx}] = {rhs_py}"]
            return [f"{name} = {rhs_py}"]
        if u.startswith("CLF1"):
            which = "RADSRTM" if "RADSRTM" in u else "RADRRTM"
            return [f"CLF1 = '{which}'"]
        raise SyntaxError("cannot transpile: " + u)

    @staticmethod
    def _split_assign(u):
        depth = 0
        for i, ch in enumerate(u):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and u[i + 1] != "=" and u[i - 1] not in "<>=/":
                return u[:i], u[i + 1 :]
        raise SyntaxError(u)

    # ------------------------------------------------------------- runtime
    def _call(self, name):
        self.run(name)

    def _open(self, unit, fname):
        self.units[unit] = open(os.path.join(self.data_dir, fname), "rb")

    def _read(self, unit, arrays):
        f = self.units[unit]
        (n,) = struct.unpack(">i", f.read(4))
        payload = f.read(n)
        (n2,) = struct.unpack(">i", f.read(4))
        assert n == n2
        vals = np.frombuffer(payload, dtype=">f8").astype(np.float64)
        pos = 0
        for arr in arrays:
            cnt = arr.a.size
            arr.a[...] = vals[pos : pos + cnt].reshape(arr.a.shape, order="F")
            pos += cnt
        assert pos == vals.size, (pos, vals.size)

"""TEST INFRASTRUCTURE — ctypes bindings for the CPU checkers.

`Port`  = oracle/libpintron_oracle.so, OUR plain-C restatement (oracle/port/dp_port.c).
`Ref`   = oracle/_ref/libref_dp.so, the UNMODIFIED reference routines (oracle/ref_dp_shim.c), present only
          where oracle/_ref was built (this container; it travels to the GPU box prebuilt).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
c_char_p, c_int, c_uint, c_long, c_size_t, c_double = C.c_char_p, C.c_int, C.c_uint, C.c_long, C.c_size_t, C.c_double


def build_port():
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)


def build_ref():
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def ops_to_rows(ops, est, gen):
    """Rebuild the two alignment rows the reference materialises from the column ops."""
    i = j = 0
    a, b = bytearray(), bytearray()
    for o in ops:
        if o == 0:
            a.append(est[i]); b.append(gen[j]); i += 1; j += 1
        elif o == 1:
            a.append(est[i]); b.append(ord('-')); i += 1
        else:
            a.append(ord('-')); b.append(gen[j]); j += 1
    return bytes(a), bytes(b)


def rows_to_ops(ra, rb):
    out = bytearray()
    for x, y in zip(ra, rb):
        out.append(2 if x == ord('-') and y != ord('-') else (1 if y == ord('-') else 0))
    return bytes(out)


class Port:
    def __init__(self):
        path = os.path.join(HERE, "libpintron_oracle.so")
        if not os.path.exists(path):
            build_port()
        self.lib = L = C.CDLL(path)
        L.po_align.restype = c_int
        L.po_edit.restype = c_uint
        L.po_suffix_cut.restype = c_uint
        L.po_prefix_cut.restype = c_uint
        L.po_seed.restype = c_long
        L.po_seed.argtypes = [c_char_p, c_long, c_char_p, c_long, c_int, c_double, C.c_void_p, c_long]
        L.po_lcs.argtypes = [c_char_p, c_long, c_char_p, c_long, C.c_void_p, C.c_void_p, C.c_void_p]

    def align(self, est, gen):
        ops = C.create_string_buffer(len(est) + len(gen) + 1)
        n = c_int()
        score = self.lib.po_align(est, len(est), gen, len(gen), ops, C.byref(n))
        return score, ops.raw[:n.value]

    def edit(self, a, b):
        return self.lib.po_edit(a, len(a), b, len(b))

    def kband(self, a, b, k):
        e = c_uint()
        ok = self.lib.po_kband(a, len(a), b, len(b), c_uint(k), C.byref(e))
        return bool(ok), e.value

    def burset(self, d, a):
        return self.lib.po_burset(C.c_char(d[0:1]), C.c_char(d[1:2]), C.c_char(a[0:1]), C.c_char(a[1:2]))

    def borders(self, p, t, max_errs, min_cut=0, max_cut=None):
        out = (c_int * 4)()
        if max_cut is None:
            max_cut = len(p)
        ok = self.lib.po_borders(p, len(p), min_cut, max_cut, t, len(t), c_uint(max_errs), out)
        return bool(ok), list(out)

    def gap(self, est, gen):
        ops = C.create_string_buffer(len(est) + len(gen) + 16)
        pos = (c_int * 5)()
        dim = self.lib.po_gap(est, len(est), gen, len(gen), ops, pos)
        return ops.raw[:dim], list(pos)

    def affix(self, est, gen):
        e, g = c_int(), c_int()
        ok = self.lib.po_affix(est, len(est), gen, len(gen), C.byref(e), C.byref(g))
        return (True, e.value, g.value) if ok else (False, 0, 0)

    def suffix_cut(self, a, b):
        x, y = c_int(), c_int()
        ed = self.lib.po_suffix_cut(a, len(a), b, len(b), C.byref(x), C.byref(y))
        return ed, x.value, y.value

    def prefix_cut(self, a, b):
        x, y = c_int(), c_int()
        ed = self.lib.po_prefix_cut(a, len(a), b, len(b), C.byref(x), C.byref(y))
        return ed, x.value, y.value

    def lcs(self, s1, s2):
        o1, o2, ln = c_long(), c_long(), c_long()
        self.lib.po_lcs(s1, len(s1), s2, len(s2), C.byref(o1), C.byref(o2), C.byref(ln))
        return ln.value, o1.value, o2.value

    def seed(self, genome, est, mfl=15, rate=0.2, cap=1 << 16):
        buf = (c_int * (3 * cap))()
        n = self.lib.po_seed(genome, len(genome), est, len(est), mfl, rate, buf, cap)
        assert n >= 0
        return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]


class Ref:
    PATH = os.path.join(HERE, "_ref", "libref_dp.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = L = C.CDLL(self.PATH)
        L.ref_edit_distance.restype = c_uint
        L.ref_edit_distance.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t]
        L.ref_compute_edit_distance.restype = c_size_t
        L.ref_compute_edit_distance.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t]
        for f in (L.ref_best_suffix_cut, L.ref_best_prefix_cut):
            f.restype = c_size_t
            f.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t, C.c_void_p, C.c_void_p]
        L.ref_refine_borders.argtypes = [c_char_p, c_size_t, c_size_t, c_size_t, c_char_p, c_size_t, c_uint,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_lcs.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_longest_affix.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t, C.c_void_p, C.c_void_p]
        L.ref_dust.restype = c_double
        L.ref_index_create.restype = C.c_void_p
        L.ref_index_create.argtypes = [c_char_p, c_int, c_double]
        L.ref_seed.restype = c_long
        L.ref_seed.argtypes = [C.c_void_p, c_char_p, c_int, C.c_void_p, c_long]

    def align(self, est, gen):
        ra = C.create_string_buffer(len(est) + len(gen) + 2)
        rb = C.create_string_buffer(len(est) + len(gen) + 2)
        dim = c_int()
        score = self.lib.ref_align(est, gen, ra, rb, C.byref(dim))
        return score, ra.raw[:dim.value], rb.raw[:dim.value]

    def edit(self, a, b):
        return self.lib.ref_edit_distance(a, len(a), b, len(b))

    def compute_edit(self, a, b):
        return self.lib.ref_compute_edit_distance(a, len(a), b, len(b))

    def kband(self, a, b, k):
        e = c_uint()
        ok = self.lib.ref_kband(a, b, c_uint(k), C.byref(e))
        return bool(ok), e.value

    def burset(self, d, a):
        return self.lib.ref_burset(d, a)

    def borders(self, p, t, max_errs, min_cut=0, max_cut=None):
        if max_cut is None:
            max_cut = len(p)
        op, o1, o2, ed = c_size_t(), c_size_t(), c_size_t(), c_uint()
        ok = self.lib.ref_refine_borders(p, len(p), min_cut, max_cut, t, len(t), max_errs,
                                         C.byref(op), C.byref(o1), C.byref(o2), C.byref(ed))
        return bool(ok), [op.value, o1.value, o2.value, ed.value]

    def gap(self, est, gen):
        ra = C.create_string_buffer(len(est) + len(gen) + 16)
        rb = C.create_string_buffer(len(est) + len(gen) + 16)
        pos = (c_int * 5)()
        dim = self.lib.ref_gap_align(est, gen, ra, rb, pos)
        return ra.raw[:dim], rb.raw[:dim], list(pos)

    def affix(self, est, gen):
        e, g = c_size_t(), c_size_t()
        ok = self.lib.ref_longest_affix(est, len(est), gen, len(gen), C.byref(e), C.byref(g))
        return (True, e.value, g.value) if ok else (False, 0, 0)

    def suffix_cut(self, a, b):
        x, y = c_size_t(), c_size_t()
        ed = self.lib.ref_best_suffix_cut(a, len(a), b, len(b), C.byref(x), C.byref(y))
        return ed, x.value, y.value

    def prefix_cut(self, a, b):
        x, y = c_size_t(), c_size_t()
        ed = self.lib.ref_best_prefix_cut(a, len(a), b, len(b), C.byref(x), C.byref(y))
        return ed, x.value, y.value

    def lcs(self, s1, s2):
        o1, o2, ln = c_size_t(), c_size_t(), c_size_t()
        self.lib.ref_lcs(s1, len(s1), s2, len(s2), C.byref(o1), C.byref(o2), C.byref(ln))
        return ln.value, o1.value, o2.value

    def index(self, genome, mfl=15, rate=0.2):
        return self.lib.ref_index_create(genome, mfl, rate)

    def seed(self, index, est, mfl=15, cap=1 << 16):
        buf = (c_int * (3 * cap))()
        n = self.lib.ref_seed(index, est, mfl, buf, cap)
        assert n >= 0
        return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]

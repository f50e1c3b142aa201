"""CPU restatement (NumPy/SciPy) of the P3DFFT++ 3D-transform path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  Nothing under p3dfft.3_b200/ imports it.

What is restated, with the reference lines each function follows (paths relative to /root/reference):
  block_dist        build/init.C:1834-1857   first P - N%P blocks get floor(N/P), the rest one more
  proc_coords       build/init.C:1654-1670   MPI_Cart_create(reorder=0): row-major coordinates
  OGrid             build/init.C:1699-1729   Pdims[i] = ProcDims[Dmap[i]], grid_id[i] = coords[Dmap[i]],
                                             Ldims / GlobStart from block_dist
  to_storage        sample/C++/test3D_r2c.C:183-187, build/exec.C:755-758   MemOrder[i] = storage rank of dim i
  transform_global  build/templ.C:232-384 (type i acts along logical dim i; R2C first, halves that dim to
                    N/2+1; C2R last) with the 1D definitions of FFTW3 (build/init.C:1146-1607):
                    unnormalised DFTs, FFTW_REDFT00/10/01/11 and FFTW_RODFT00/10/01/11, "_COMPLEX" r2r =
                    re and im transformed separately (init.C:1179-1188)
  deriv_factor      build/exec.C:228-287     i*k below g/2, 0 at g/2, i*(k-g) above; g=(n-1)*2 after R2C
  compute_deriv     build/deriv.C:85-185     stand-alone derivative incl. its inverse-permutation choice of
                                             the storage dimension (deriv.C:90-94)

The 1D arithmetic lives in FFTW3, a third-party dependency that is NOT vendored in the reference (version
unpinned, configure.ac:129-153).  Its published definitions are restated through numpy.fft (pocketfft) and
scipy.fft.dct/dst(norm=None); oracle/cfft/ holds an independent plain-C restatement used by oracle/_ref.

Pinning: tests/test_oracle.py checks this module against (a) the known-answer spectra of the reference's
own samples (test3D_r2c.C:281-331, test1D_cos.C:254-294, test1D_sin.C, test_deriv2.C:354-431) and
(b) golden per-rank arrays produced by the reference's unmodified host code (oracle/_ref, built from
/root/reference by oracle/Makefile) and committed under tests/golden/.
"""
import numpy as np

try:
    import scipy.fft as _sfft
except Exception:  # pragma: no cover - scipy is part of the image
    _sfft = None

# ---------------------------------------------------------------------------------------------- type table
# IDs in registration order (build/init.C:121-753): name -> (kind, dt_in, dt_out, prec)
TYPE_TABLE = []
for _n, _k, _d1, _d2 in (("EMPTY_TYPE_SINGLE", "empty", 1, 1), ("EMPTY_TYPE_DOUBLE", "empty", 1, 1),
                         ("EMPTY_TYPE_SINGLE_COMPLEX", "empty", 2, 2), ("EMPTY_TYPE_DOUBLE_COMPLEX", "empty", 2, 2)):
    TYPE_TABLE.append((_n, _k, _d1, _d2, 4 if "SINGLE" in _n else 8))
TYPE_TABLE += [("R2CFFT_S", "r2c", 1, 2, 4), ("R2CFFT_D", "r2c", 1, 2, 8), ("C2RFFT_S", "c2r", 2, 1, 4),
               ("C2RFFT_D", "c2r", 2, 1, 8), ("CFFT_FORWARD_S", "fwd", 2, 2, 4), ("CFFT_FORWARD_D", "fwd", 2, 2, 8),
               ("CFFT_BACKWARD_S", "bwd", 2, 2, 4), ("CFFT_BACKWARD_D", "bwd", 2, 2, 8)]
for _num in (1, 2, 3, 4):
    for _fam in ("DCT", "DST"):
        for _var, _dt, _pr in (("REAL_S", 1, 4), ("REAL_D", 1, 8), ("COMPLEX_S", 2, 4), ("COMPLEX_D", 2, 8)):
            TYPE_TABLE.append((f"{_fam}{_num}_{_var}", f"{_fam.lower()}{_num}", _dt, _dt, _pr))
TYPE_ID = {t[0]: i for i, t in enumerate(TYPE_TABLE)}


# The reference registers its four DCT4 IDs with the DCT-I planner (build/init.C:640,652,664,676: plan_dct1_*), so
# a "DCT4" request executes FFTW_REDFT00.  Pinned by the golden cases t1d_DCT4_REAL_D_* (reference host code).
REFERENCE_DCT4_IS_DCT1 = True


def type_info(t):
    """(kind, dt_in, dt_out, prec) of a type given by name or ID, with the reference's DCT4 registration quirk"""
    rec = TYPE_TABLE[TYPE_ID[t]] if isinstance(t, str) else TYPE_TABLE[int(t)]
    kind = rec[1]
    if kind == "dct4" and REFERENCE_DCT4_IS_DCT1:
        kind = "dct1"
    return kind, rec[2], rec[3], rec[4]


# ---------------------------------------------------------------------------------------------- geometry
def block_dist(n, p):
    """start[], size[] of the p blocks of a dimension of n points (init.C:1834-1857)"""
    base, nlow = n // p, p - n % p
    size = [base if b < nlow else base + 1 for b in range(p)]
    start = [sum(size[:b]) for b in range(p)]
    return start, size


def proc_coords(rank, procdims):
    """Cartesian coordinates of a rank, row-major, last dimension fastest (init.C:1654-1670)"""
    c = [0, 0, 0]
    r = rank
    for i in (2, 1, 0):
        c[i] = r % procdims[i]
        r //= procdims[i]
    return c


def inv_mo(mo):
    imo = [0, 0, 0]
    for i in range(3):
        imo[mo[i]] = i
    return imo


class OGrid:
    """DataGrid of one rank (init.C:1699-1729)"""

    def __init__(self, gdims, dmap, mem_order, procdims, rank, dim_conj_sym=-1):
        self.Gdims = list(gdims)
        self.Dmap = list(dmap)
        self.MemOrder = list(mem_order)
        self.ProcDims = list(procdims)
        self.rank = rank
        self.dim_conj_sym = dim_conj_sym
        coords = proc_coords(rank, procdims)
        self.Pdims = [procdims[dmap[i]] for i in range(3)]
        self.grid_id = [coords[dmap[i]] for i in range(3)]
        self.Ldims, self.GlobStart = [], []
        for i in range(3):
            st, sz = block_dist(gdims[i], self.Pdims[i])
            self.Ldims.append(sz[self.grid_id[i]])
            self.GlobStart.append(st[self.grid_id[i]])

    def slices(self):
        return tuple(slice(self.GlobStart[i], self.GlobStart[i] + self.Ldims[i]) for i in range(3))

    def storage_shape(self):
        """numpy shape (slowest..fastest) of the local array"""
        imo = inv_mo(self.MemOrder)
        return tuple(self.Ldims[imo[r]] for r in (2, 1, 0))


def to_storage(block, mem_order):
    """logical block [i0,i1,i2] -> C-contiguous local array whose LAST numpy axis is storage rank 0"""
    imo = inv_mo(mem_order)
    return np.ascontiguousarray(np.transpose(block, (imo[2], imo[1], imo[0])))


def from_storage(local, mem_order):
    """inverse of to_storage: local array (numpy axes = storage ranks 2,1,0) -> logical block"""
    imo = inv_mo(mem_order)
    # numpy axis a holds logical dim imo[2-a]
    perm = [0, 0, 0]
    for a in range(3):
        perm[imo[2 - a]] = a
    return np.transpose(local, perm)


def local_of(G, grid):
    """this rank's local array of the global logical array G"""
    return to_storage(G[grid.slices()], grid.MemOrder)


def assemble(locals_, grids, dtype=None):
    """global logical array from all ranks' local arrays"""
    g0 = grids[0]
    G = np.zeros(g0.Gdims, dtype=dtype or locals_[0].dtype)
    for loc, gr in zip(locals_, grids):
        G[gr.slices()] = from_storage(np.asarray(loc).reshape(gr.storage_shape()), gr.MemOrder)
    return G


# ---------------------------------------------------------------------------------------------- 1D transforms
def _r2r(x, kind, axis):
    fam, num = kind[:3], int(kind[3])
    f = _sfft.dct if fam == "dct" else _sfft.dst

    def one(a):
        return f(a, type=num, axis=axis, norm=None)

    if np.iscomplexobj(x):  # init.C:1179-1188: real and imaginary parts separately
        return one(x.real) + 1j * one(x.imag)
    return one(x)


def transform_1d(x, kind, axis, n_real=None):
    """unnormalised 1D transform of FFTW's definition along `axis` (double precision arithmetic)"""
    if kind == "empty":
        return x
    if kind == "fwd":
        return np.fft.fft(x, axis=axis)
    if kind == "bwd":
        return np.fft.ifft(x, axis=axis) * x.shape[axis]
    if kind == "r2c":
        return np.fft.rfft(x, axis=axis)
    if kind == "c2r":
        n = n_real if n_real is not None else (x.shape[axis] - 1) * 2
        return np.fft.irfft(x, n=n, axis=axis) * n
    return _r2r(x, kind, axis)


def transform_order(types):
    """order in which the three 1D transforms are applied: R2C first, C2R last (templ.C:639-655)"""
    kinds = [type_info(t)[0] for t in types]
    first = [i for i in range(3) if kinds[i] == "r2c"]
    last = [i for i in range(3) if kinds[i] == "c2r"]
    mid = [i for i in range(3) if i not in first and i not in last]
    return first + mid + last


def transform_global(G, types, gdims_out=None, deriv_dim=-1):
    """3D transform of the global logical array G[i0,i1,i2]; type i acts along dim i.
    deriv_dim >= 0: also apply the spectral derivative right after the transform of that dim (exec.C:175-199)."""
    A = np.asarray(G)
    A = A.astype(np.complex128 if np.iscomplexobj(A) else np.float64)
    for d in transform_order(types):
        kind = type_info(types[d])[0]
        n_real = gdims_out[d] if (kind == "c2r" and gdims_out is not None) else None
        A = transform_1d(A, kind, d, n_real)
        if d == deriv_dim and kind != "empty":
            n = A.shape[d]
            g = (n - 1) * 2 if kind == "r2c" else n
            shp = [1, 1, 1]
            shp[d] = n
            A = A * deriv_factor(np.arange(n), g).reshape(shp)
    return A


# ---------------------------------------------------------------------------------------------- derivative
def deriv_factor(k, g):
    """i*kappa for global wavenumber index k and full length g (exec.C:228-287, deriv.C:100-185)"""
    k = np.asarray(k)
    mid = g // 2
    kap = np.where(k < mid, k, np.where(k == mid, 0, k - g))
    return 1j * kap.astype(np.float64)


def compute_deriv_local(local, grid, idir, mode="reference"):
    """stand-alone derivative of a rank's local complex array (deriv.C:85-185).  The reference picks the
    storage dimension by the INVERSE permutation (ldir = i with MemOrder[i] == idir, deriv.C:90-94);
    mode="memorder" uses MemOrder[idir] (what the comment in the source intends)."""
    mo = grid.MemOrder
    ldir = [i for i in range(3) if mo[i] == idir][0] if mode == "reference" else mo[idir]
    g = (grid.Gdims[idir] - 1) * 2 if grid.dim_conj_sym == idir else grid.Gdims[idir]
    a = np.asarray(local).reshape(grid.storage_shape())
    n = a.shape[2 - ldir]
    fac = deriv_factor(np.arange(n) + grid.GlobStart[idir], g)
    shp = [1, 1, 1]
    shp[2 - ldir] = n
    return a * fac.reshape(shp)


# ---------------------------------------------------------------------------------------------- synthetic fields
def random_field(gdims, complex_=False, key=20240):
    """globally indexed random field: logical array [i0,i1,i2], double precision"""
    def draw(k):
        rng = np.random.Generator(np.random.Philox(key=k))
        return rng.standard_normal((gdims[2], gdims[1], gdims[0])).transpose(2, 1, 0)
    a = draw(key)
    if complex_:
        a = a + 1j * draw(key + 1)
    return np.ascontiguousarray(a)


def sine_field(gdims):
    """sin(2 pi x/Nx) sin(2 pi y/Ny) sin(2 pi z/Nz)  (sample/C++/test3D_r2c.C:343-369)"""
    s = [np.sin(2 * np.pi * np.arange(n) / n) for n in gdims]
    return s[0][:, None, None] * s[1][None, :, None] * s[2][None, None, :]


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))

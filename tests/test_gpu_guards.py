"""Out-of-bounds guard tests.  compute-sanitizer is closed on this GPU pool (gpurun refuses it: "runs under it have left GPUs
needing a reset", profiles/r02_sanitizer_closed.txt), so every buffer the round-2 kernels write -- digit planes, the double-buffered
predict / wave workspaces, outputs, compaction and Saltelli scratch -- is allocated here EXACTLY at the size the C ABI asks for,
between two 64 KB canary zones, at ragged sizes (points not a multiple of the chunk, train size not a multiple of 128); any write
outside the requested bytes flips a canary.  Results are checked against the FP64 path at the same time."""
import ctypes as C

import numpy as np
import pytest
import torch

from gperks_b200 import _native as nv

pytestmark = pytest.mark.gpu
F64 = torch.float64
PAD = 65536


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


class Guarded:
    """nbytes of device memory between two canary zones; .ptr is 256-byte aligned."""

    def __init__(self, nbytes, dev, fill=0xA5):
        self.n = int(nbytes)
        self.raw = torch.full((self.n + 2 * PAD + 256,), fill, dtype=torch.uint8, device=dev)
        base = self.raw.data_ptr()
        self.off = (-(base + PAD)) % 256 + PAD
        self.ptr = base + self.off
        self.fill = fill

    def view(self, dtype, count):
        return self.raw[self.off: self.off + self.n].view(dtype)[:count]

    def intact(self):
        lo = self.raw[: self.off]
        hi = self.raw[self.off + self.n:]
        return bool((lo == self.fill).all()) and bool((hi == self.fill).all())


def _engine(dev, N, D, noise=1e-2, kind=3, seed=0):
    from gperks_b200.engine import ExactGPEngine
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, D, dtype=F64, generator=g).to(dev)
    y = torch.randn(N, dtype=F64, generator=g).to(dev)
    theta = torch.cat([torch.full((D,), 1.7, dtype=F64), torch.tensor([1.1, noise], dtype=F64)]).to(dev)
    eng = ExactGPEngine(X, kind)
    eng.factorize(theta, y)
    return eng


@pytest.mark.parametrize("S,bits", [(5, 8), (6, 8), (6, 7)])
@pytest.mark.parametrize("overlap", [True, False])
def test_predict_i8_stays_inside_its_buffers(dev, S, bits, overlap):
    lib = nv.lib()
    N, D, M, Mc = 300, 3, 1000, 256                      # Np = 384, chunks 256 + 256 + 256 + 232
    eng = _engine(dev, N, D)
    eng._ensure_i8(S, bits)
    Xs = torch.rand(M, D, dtype=F64, device=dev)
    m64, s64 = eng.predict(Xs, precision="fp64")
    ws = Guarded(lib.gpx_predict_i8_workspace_bytes(eng.Np, Mc, S), dev)
    om, os_ = Guarded(M * 8, dev), Guarded(M * 8, dev)
    mma = torch.cuda.Stream(priority=-1) if overlap else None
    prior = torch.zeros(M, dtype=F64, device=dev)
    p = nv.ptr
    nv.check(lib.gpx_predict_meanvar_i8(p(Xs), M, p(eng.X), N, D, p(eng.theta), None, None, eng.kind, p(eng.S_i8), p(eng.S_i8_inv_scale),
                                        S, bits, eng.Np, p(eng.alpha), p(prior), None, 0, 0, None, None, 0.0, 1.0, 1e-10, om.ptr, os_.ptr,
                                        ws.ptr, ws.n, Mc, nv.stream_ptr(), mma.cuda_stream if mma else None), "predict_i8")
    torch.cuda.synchronize()
    assert ws.intact() and om.intact() and os_.intact()
    assert torch.equal(om.view(F64, M), m64)
    np.testing.assert_allclose(os_.view(F64, M).cpu().numpy(), s64.cpu().numpy(), rtol=1e-6)
    # the plane writers on their own
    planes = Guarded(S * Mc * eng.Np, dev)
    mp = Guarded(eng.nb * Mc * 8, dev)
    nv.check(lib.gpx_kmat_cross_i8(p(Xs), 200, p(eng.X), N, D, p(eng.theta), None, None, eng.kind, p(eng.alpha), planes.ptr, Mc, S, bits,
                                   eng.Np, mp.ptr, Mc, nv.stream_ptr()), "builder")
    Ls = Guarded(S * eng.Np * eng.Np, dev)
    isc = Guarded(eng.Np * 8, dev)
    nv.check(lib.gpx_slice_i8(p(eng.S), eng.Np, eng.Np, eng.Np, Ls.ptr, eng.Np, S, bits, 1, isc.ptr, None, nv.stream_ptr()), "slice")
    nt = lib.gpx_i8_row_tiles(eng.Np)
    q = Guarded(nt * Mc * 8, dev)
    sched = torch.zeros(2, dtype=torch.int32, device=dev)
    nv.check(lib.gpx_trmm_sumsq_i8(Ls.ptr, isc.ptr, eng.Np, planes.ptr, Mc, S, bits, p(eng.theta) + 8 * D, q.ptr, p(sched), nv.stream_ptr()), "mma")
    torch.cuda.synchronize()
    assert planes.intact() and mp.intact() and Ls.intact() and isc.intact() and q.intact()


def test_wave_compaction_saltelli_and_mean_stay_inside_their_buffers(dev):
    lib = nv.lib()
    D, M, Mc = 3, 700, 256
    engs = [_engine(dev, n, D, noise=nz, kind=k, seed=n) for n, nz, k in ((300, 1e-2, 3), (520, 1e-3, 0))]
    descs = []
    for eng in engs:
        eng._ensure_i8(5, 8)
        descs.append(nv.I8EmulatorDesc(X=nv.ptr(eng.X), theta=nv.ptr(eng.theta), xa=None, xr=None, alpha=nv.ptr(eng.alpha), Ls=nv.ptr(eng.S_i8),
                                       inv_scale=nv.ptr(eng.S_i8_inv_scale), feat_idx=None, weights=None, bias=None, y_mu=0.0, y_sigma=1.0,
                                       min_var=1e-10, N=eng.N, Np=eng.Np, kind=eng.kind, nslices=5, bits=8, n_feat=0, degree=0, reserved=0))
    arr = (nv.I8EmulatorDesc * 2)(*descs)
    Xs = torch.rand(M, D, dtype=F64, device=dev)
    z = torch.tensor([0.1, -0.2], dtype=F64, device=dev)
    v = torch.tensor([0.05, 0.02], dtype=F64, device=dev)
    ws = Guarded(lib.gpx_wave_i8_workspace_bytes(arr, 2, Mc), dev)
    I, PV = Guarded(M * 8, dev), Guarded(M * 8, dev)
    mma = torch.cuda.Stream(priority=-1)
    nv.check(lib.gpx_wave_i8(arr, 2, D, nv.ptr(Xs), M, nv.ptr(z), nv.ptr(v), 1, I.ptr, PV.ptr, None, None, ws.ptr, ws.n, Mc, nv.stream_ptr(),
                             mma.cuda_stream), "wave")
    torch.cuda.synchronize()
    assert ws.intact() and I.intact() and PV.intact()
    means, stds = zip(*[e.predict(Xs, precision="int8x5w") for e in engs])
    Iref = torch.stack([((m - z[i]) ** 2 / (s * s + v[i])).sqrt() for i, (m, s) in enumerate(zip(means, stds))]).max(0).values
    np.testing.assert_allclose(I.view(F64, M).cpu().numpy(), Iref.cpu().numpy(), rtol=1e-12)
    # compaction
    idx, cnt = Guarded(M * 8, dev), Guarded(8, dev)
    cws = Guarded(lib.gpx_compact_below_workspace_bytes(M), dev)
    nv.check(lib.gpx_compact_below(I.ptr, M, float(Iref.median()), idx.ptr, cnt.ptr, cws.ptr, cws.n, nv.stream_ptr()), "compact")
    # Saltelli sums and the mean-only predict
    n, d, K = 4097, D, 5
    JC = lib.gpx_saltelli_chunk_rows()
    nc = (n + JC - 1) // JC
    f = torch.randn(n * (d + 2), dtype=F64, device=dev)
    part, sums = Guarded(nc * K * (4 + 3 * d) * 8, dev), Guarded(K * (4 + 3 * d) * 8, dev)
    nv.check(lib.gpx_saltelli_partial(nv.ptr(f), nv.ptr(f) + 8 * n, nv.ptr(f) + 16 * n, nv.ptr(f), nv.ptr(f) + 8 * n, nv.ptr(f) + 16 * n, n, d, 0, n,
                                      K, 3, part.ptr, nv.stream_ptr()), "saltelli")
    nv.check(lib.gpx_saltelli_reduce(part.ptr, nc, d, K, sums.ptr, nv.stream_ptr()), "reduce")
    om = Guarded(M * 8, dev)
    e0 = engs[0]
    nv.check(lib.gpx_predict_mean(nv.ptr(Xs), M, nv.ptr(e0.X), e0.N, D, nv.ptr(e0.theta), None, None, e0.kind, nv.ptr(e0.alpha), e0.Np, None, None, 0, 0,
                                  None, None, 0.0, 1.0, om.ptr, nv.stream_ptr()), "mean")
    torch.cuda.synchronize()
    assert idx.intact() and cnt.intact() and cws.intact() and part.intact() and sums.intact() and om.intact()
    assert torch.equal(om.view(F64, M), means[0])
    assert C.sizeof(nv.I8EmulatorDesc) == 136


def test_training_mode_integer_kernels_stay_inside_their_buffers(dev):
    """gpx_potrf_i8 (sliced panels in alternating plane buffers, accumulate epilogue into K), gpx_predict_covar_i8 (V -> W, V^T -> Vt
    through the mirror store, accumulate into C with its own leading dimension) and gpx_i8_digit_bound, every buffer exactly sized
    between canaries, at a train size whose last block column is ragged and whose outer blocks leave a ragged rest."""
    lib = nv.lib()
    p = nv.ptr
    N, D = 3150, 4                                      # Np = 3200: 25 block columns -> one outer block of the two-level schedule
    eng = _engine(dev, N, D)
    Np, nb = eng.Np, eng.Np // 128
    assert Np == 3200 and eng._potrf_i8_capable
    L64 = eng.chol().clone()
    st = nv.stream_ptr()
    K = Guarded(Np * Np * 8, dev)
    S, DT, logdet, info = Guarded(Np * Np * 8, dev), Guarded(nb * 128 * 128 * 8, dev), Guarded(nb * 8, dev), Guarded(4, dev)
    ws = Guarded(lib.gpx_potrf_i8_workspace_bytes(Np), dev)
    nv.check(lib.gpx_kmat_sym(p(eng.X), N, D, p(eng.theta), eng.kind, K.ptr, Np, st), "kmat")
    nv.check(lib.gpx_potrf_i8(K.ptr, Np, N, S.ptr, DT.ptr, logdet.ptr, info.ptr, ws.ptr, ws.n, st), "potrf_i8")
    torch.cuda.synchronize()
    assert K.intact() and S.intact() and DT.intact() and logdet.intact() and info.intact() and ws.intact()
    assert int(info.view(torch.int32, 1).item()) == 0
    L8 = torch.tril(K.view(F64, Np * Np).view(Np, Np)[:N, :N])
    err = float((L8 - L64).abs().max() / L64.abs().max())
    assert 0 < err < 1e-10, err
    # dense predictive covariance of M validation points on the integer GEMM
    M = 1100
    Mp = nv.padded(M)
    Xs = torch.rand(M, D, dtype=F64, device=dev)
    Ks, Vt, Cc, W = Guarded(Mp * Np * 8, dev), Guarded(Mp * Np * 8, dev), Guarded(Mp * Mp * 8, dev), Guarded(Np * Np * 8, dev)
    mp1, mp2 = Guarded(nb * Mp * 8, dev), Guarded((Mp // 128) * Mp * 8, dev)
    zeros = torch.zeros(Mp, dtype=F64, device=dev)
    ws2 = Guarded(lib.gpx_trtri_i8_workspace_bytes(Np), dev)
    nv.check(lib.gpx_kmat_cross(p(Xs), M, p(eng.X), N, D, p(eng.theta), None, None, eng.kind, p(eng.alpha), Ks.ptr, Np, mp1.ptr, Mp, st), "cross")
    nv.check(lib.gpx_kmat_cross(p(Xs), M, p(Xs), M, D, p(eng.theta_latent), None, None, eng.kind, p(zeros), Cc.ptr, Mp, mp2.ptr, Mp, st), "kss")
    nv.check(lib.gpx_predict_covar_i8(p(eng.S), Np, Ks.ptr, Mp, Vt.ptr, Cc.ptr, W.ptr, 8, ws2.ptr, ws2.n, st), "covar_i8")
    torch.cuda.synchronize()
    assert Ks.intact() and Vt.intact() and Cc.intact() and W.intact() and mp1.intact() and mp2.intact() and ws2.intact()
    _, cref = eng.predict_covar(Xs, torch.zeros(M, dtype=F64, device=dev), noisy=False)
    got = torch.tril(Cc.view(F64, Mp * Mp).view(Mp, Mp)[:M, :M])
    assert float((got - torch.tril(cref)).abs().max()) < 1e-9 * float(eng.theta[D])
    # the digit bound of the sliced L^-1
    eng._ensure_i8(5, 8)
    out = Guarded(8, dev)
    nv.check(lib.gpx_i8_digit_bound(p(eng.S_i8), 5, Np, out.ptr, st), "digit_bound")
    torch.cuda.synchronize()
    assert out.intact()
    rows = torch.arange(Np, device=dev)
    live = torch.arange(Np, device=dev).view(1, -1) < ((rows // 128 + 1) * 128).view(-1, 1)     # the k range the MMA kernel reads per row
    ref = (eng.S_i8.to(torch.int32).abs() * live.unsqueeze(0)).sum(dim=(0, 2)).max()
    assert int(out.view(torch.int64, 1).item()) == int(ref)

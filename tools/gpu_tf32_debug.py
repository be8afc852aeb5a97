import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gperks_b200 import _native as nv
lib = nv.lib()
dev = torch.device("cuda:0")
F32, F64 = torch.float32, torch.float64

def run(Khi, Klo, Shi, Slo):
    Mc, Np = Khi.shape
    nt = lib.gpx_tf32_row_tiles(Np)
    q = torch.full((nt, Mc), -1.0, dtype=F64, device=dev)
    rc = lib.gpx_trmm_sumsq_tf32(nv.ptr(Shi), nv.ptr(Slo), Np, nv.ptr(Khi), nv.ptr(Klo), Mc, nv.ptr(q), 128, nv.stream_ptr())
    torch.cuda.synchronize()
    assert rc == 0, rc
    return q

def ref(Khi, Klo, Shi, Slo):
    K = Khi.double() + Klo.double(); S = torch.tril(Shi.double() + Slo.double())
    Np = S.shape[0]; nt = (Np + 255) // 256
    V = K @ S.T
    # kernel computes full 256-blocks with k < 256(it+1): rows beyond diag inside tile use what's stored (we pass tril-masked S)
    out = []
    for it in range(nt):
        out.append((V[:, it*256:(it+1)*256] ** 2).sum(1))
    return torch.stack(out)

torch.manual_seed(0)
for Np, Mc in [(256, 256), (512, 512), (1024, 256)]:
    z = lambda *s: torch.zeros(*s, dtype=F32, device=dev)
    # test 1: K = ones, S = I
    K1 = torch.ones(Mc, Np, dtype=F32, device=dev); S1 = torch.eye(Np, dtype=F32, device=dev)
    q = run(K1, z(Mc, Np), S1, z(Np, Np)); r = ref(K1, z(Mc, Np), S1, z(Np, Np))
    print(Np, Mc, "T1 ones x I   :", q[:, :4].tolist(), "ref", r[:, :4].tolist())
    # test 2: K[m][k] = k % 7, S = I
    K2 = (torch.arange(Np, device=dev) % 7).float().repeat(Mc, 1).contiguous()
    q = run(K2, z(Mc, Np), S1, z(Np, Np)); r = ref(K2, z(Mc, Np), S1, z(Np, Np))
    print(Np, Mc, "T2 k%7 x I    :", q[:, :3].tolist(), "ref", r[:, :3].tolist())
    # test 3: K[m][k] = m % 5 + 1, S = I
    K3 = ((torch.arange(Mc, device=dev) % 5) + 1).float().unsqueeze(1).repeat(1, Np).contiguous()
    q = run(K3, z(Mc, Np), S1, z(Np, Np)); r = ref(K3, z(Mc, Np), S1, z(Np, Np))
    print(Np, Mc, "T3 m%5+1 x I  :", q[0, :6].tolist(), "ref", r[0, :6].tolist())
    # test 4: random tf32-exact small ints, lower-triangular S
    K4 = torch.randint(-3, 4, (Mc, Np), device=dev).float(); S4 = torch.tril(torch.randint(-2, 3, (Np, Np), device=dev).float())
    q = run(K4, z(Mc, Np), S4, z(Np, Np)); r = ref(K4, z(Mc, Np), S4, z(Np, Np))
    print(Np, Mc, "T4 randint    : maxabs diff", float((q - r).abs().max()), "max ref", float(r.max()))
    # test 5: lo terms
    K5l = torch.randint(-3, 4, (Mc, Np), device=dev).float() * 2**-12; S5l = torch.tril(torch.randint(-2, 3, (Np, Np), device=dev).float()) * 2**-12
    q = run(K4, K5l, S4, S5l); r = ref(K4, K5l, S4, S5l)
    print(Np, Mc, "T5 with lo    : rel diff", float(((q - r).abs() / r.clamp_min(1)).max()))

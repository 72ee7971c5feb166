import sys, torch
sys.path.insert(0, "/root/repo")
import __graft_entry__ as e
pkg = e.load_package(); B = pkg.binding
ctx = pkg.Context(0); ctx.use_torch_stream()
N, n, t = 1 << 26, 32, 15
pl = torch.empty((t + 1, N), dtype=torch.int64, device="cuda"); sh = torch.empty((n, N), dtype=torch.int64, device="cuda")
ctx.random_dev(61, "p", 0, (t + 1) * N, pl)
for _ in range(4): ctx.shamir_share_coeffs_dev(61, pl, N, t, n, sh, B.PARTY_MAJOR)
torch.cuda.synchronize()

import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from checkerpose_b200 import ops
torch.set_grad_enabled(False)
def r(t): return t.to(torch.bfloat16).float()
g = torch.Generator().manual_seed(6)
B, N, Co, K = 3, 200, 256, 20
z = r(torch.randn(B, N, 2 * Co, generator=g))
idx = torch.randint(0, N, (2, N, K), generator=g).int()
sel = torch.tensor([1, 0, 1]).int()
w = r(torch.randn(512, Co, generator=g) / 16)
def expect(sel_):
    gat = z[:, :, :Co][torch.arange(B)[:, None, None], idx[sel_.long()].long()]
    return r(torch.nn.functional.leaky_relu(gat.max(dim=2)[0] + z[:, :, Co:], 0.2))
a_bf = expect(sel)
zc = z.cuda().to(torch.bfloat16); ic = idx.cuda(); sc = sel.cuda()
wp = ops.pack_weight(w.cuda())
out = torch.empty((B, N, 512), dtype=torch.bfloat16, device="cuda")
a_out = torch.empty((B, N, Co), dtype=torch.bfloat16, device="cuda")
L = [ops.chain_layer(wp, None, Co, 512, False, 0.0)]
ops.chain_fwd(prologue=ops.PRO_AGG, B=B, N=N, z=zc, idx32=ic, graph_sel=sc, agg_slope=0.2, a_out=a_out, layers=L, out=out, out_mode=ops.OUT_BF16)
torch.cuda.synchronize()
got = a_out.cpu().float()
mis = got != a_bf
print("chain AGG mismatches:", int(mis.sum()), "of", mis.numel())
print("per batch:", mis.sum((1, 2)).tolist())
print("rows with mismatch (b=0):", mis[0].any(1).nonzero().flatten()[:20].tolist())
print("per 16B-chunk:", mis.view(B, N, 32, 8).sum((0, 1, 3)).tolist())
y = ops.edge_aggregate(zc, ic, sc, 0.2).cpu().float()
m2 = y != a_bf
print("SIMT aggregate mismatches:", int(m2.sum()), "per batch", m2.sum((1, 2)).tolist())
print("chain vs SIMT mismatches:", int((y != got).sum()))
for alt in ([0, 0, 0], [1, 1, 1], [0, 1, 0]):
    e = expect(torch.tensor(alt))
    print("alt sel", alt, "chain mism", int((got != e).sum()), "simt mism", int((y != e).sum()))
ref = a_bf.double() @ w.double().t()
print("gemm out close:", torch.allclose(out.cpu().double(), got.double() @ w.double().t(), rtol=1e-2, atol=1e-2))

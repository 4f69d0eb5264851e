"""CPU emulation of the reduced-precision head (rounding at the points where the CUDA kernels round) to
budget the bf16 / fp16 error against the fp32 oracle.  Diagnostic only; not part of the product or tests."""
import os
import sys

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from checkerpose_b200 import synthetic as syn  # noqa: E402
from oracle import checkerpose_oracle as orc  # noqa: E402

torch.set_grad_enabled(False)


def make_round(dt, on=True):
    if not on or dt is None:
        return lambda t: t
    return lambda t: t.to(dt).float()


def fold(sd, p, eps=1e-5):
    w = sd[p + "conv.0.weight"][:, :, 0, 0]
    Co, C2 = w.shape
    C = C2 // 2
    s = sd[p + "conv.1.weight"] / torch.sqrt(sd[p + "conv.1.running_var"] + eps)
    t = sd[p + "conv.1.bias"] - s * sd[p + "conv.1.running_mean"]
    W1, W2 = w[:, :C], w[:, C:]
    return torch.cat([s[:, None] * W1, s[:, None] * (W2 - W1)], 0), torch.cat([torch.zeros(Co), t]), Co


def edgeconv(x, idx, sd, p, slope, ra, rw, rz):
    """x (B,N,C) already rounded as A operand."""
    wf, bf, Co = fold(sd, p)
    z = rz(x @ rw(wf).t() + bf)
    P, Q = z[..., :Co], z[..., Co:]
    g = P[:, idx[0]]                    # (B,N,K,Co)
    return ra(F.leaky_relu(g.max(2)[0] + Q, slope))


def lin(x, sd, p, slope, ra, rw, act=True):
    y = x @ rw(sd[p + ".weight"]).t() + sd[p + ".bias"]
    return ra(F.leaky_relu(y, slope)) if act else y


def conv_block(x, sd, prefix, convT, ri, rw):
    def bnfold(wkey, bnkey, transposed):
        w = sd[wkey]
        s = sd[bnkey + ".weight"] / torch.sqrt(sd[bnkey + ".running_var"] + 1e-5)
        sh = sd[bnkey + ".bias"] - s * sd[bnkey + ".running_mean"]
        w = w * (s.view(1, -1, 1, 1) if transposed else s.view(-1, 1, 1, 1))
        return rw(w), ri(sh)
    if convT:
        w, b = bnfold(prefix + "0.weight", prefix + "1", True)
        x = ri(F.relu(ri(F.conv_transpose2d(x, w, b, stride=2, padding=1, output_padding=1))))
        for a, c in (("3", "4"), ("6", "7")):
            w, b = bnfold(prefix + a + ".weight", prefix + c, False)
            x = ri(F.relu(ri(F.conv2d(x, w, b, padding=1))))
    else:
        x = ri(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))
        for a, c in (("1", "2"), ("4", "5")):
            w, b = bnfold(prefix + a + ".weight", prefix + c, False)
            x = ri(F.relu(ri(F.conv2d(x, w, b, padding=1))))
    return x


def head(feats, sd, idx, N, dt, weights_exact=False, img_exact=False, z_exact=False):
    ra = make_round(dt)                       # activations (A operands, stored features)
    rw = make_round(dt, not weights_exact)    # weights
    rz = make_round(dt, not z_exact)          # [P|Q] table
    ri = make_round(dt, not img_exact)        # image branch tensors
    riw = make_round(dt, not img_exact)
    B = feats[0].shape[0]
    f3 = ri(feats[-1])
    x0 = ri(F.conv2d(f3, riw(sd["init_net.conv1x1.weight"]), ri(sd["init_net.conv1x1.bias"])))
    x = ra(x0.reshape(B, N, 64))
    for i in range(2):
        x = edgeconv(x, idx, sd, f"init_net.pre_query_block.{i}.", 0.2, ra, rw, rz)
    logits = x @ rw(sd["init_net.mlp.weight"]).t() + sd["init_net.mlp.bias"]
    roi, xb, yb = logits[..., 0], logits[..., 1:4], logits[..., 4:7]
    mask = (roi > 0).float()
    xid = ((xb > 0).long() * torch.tensor([4, 2, 1])).sum(-1)
    yid = ((yb > 0).long() * torch.tensor([4, 2, 1])).sum(-1)
    g = x
    img = f3
    xbits, ybits = [xb], [yb]
    for i in range(3):
        if i > 0:
            img = torch.cat([img, ri(feats[-i - 1])], 1)
        img = conv_block(img, sd, f"up_net.{i}.", i == 0, ri, riw)
        p = f"refine_net.{i}."
        patches = ri(F.conv2d(img, riw(sd[p + "local_feat_ext_block.patch_generator.weight"]),
                              ri(sd[p + "local_feat_ext_block.patch_generator.bias"]), padding=1))
        pn = patches.permute(0, 2, 3, 1)
        bi = torch.arange(B).view(B, 1).expand(-1, N)
        taps = torch.cat([pn[bi, 2 * yid + dy, 2 * xid + dx] for dx, dy in ((0, 0), (0, 2), (2, 0), (2, 2))], 2)
        h = torch.cat([taps * mask[..., None], g], 2)
        h = lin(h, sd, p + "pre_graph_module.0", 0.01, ra, rw)
        h = lin(h, sd, p + "pre_graph_module.2", 0.01, ra, rw)
        for j in range(3):
            h = edgeconv(h, idx, sd, f"{p}pre_query_block.{j}.", 0.2, ra, rw, rz)
        g = h
        t = lin(h, sd, p + "query_block.mlps.0", 0.01, ra, rw)
        t = lin(t, sd, p + "query_block.mlps.2", 0.01, ra, rw)
        nb = lin(t, sd, p + "query_block.mlps.4", 0.01, ra, rw, act=False)
        xbits.append(nb[..., 0:1])
        ybits.append(nb[..., 1:2])
        xid = xid * 2 + (nb[..., 0] > 0).long()
        yid = yid * 2 + (nb[..., 1] > 0).long()
    return roi, torch.cat(xbits, -1), torch.cat(ybits, -1), xid, yid


def main():
    N, B = int(os.environ.get("N", 512)), 2
    g = torch.Generator().manual_seed(2024)
    p3d = syn.p3d_normed_tensor(syn.load_fps_xyz("lmo", 1, N))
    sd = syn.synthetic_state_dict(syn.head_param_spec(N), g)
    feats = syn.synthetic_features(B, g)
    idx = orc.knn(p3d, 20)
    ref = orc.pose_head(feats, sd, idx, [idx] * 3, N)
    rroi, rxb, ryb = ref[0][:, 0], ref[1].permute(0, 2, 1), ref[2].permute(0, 2, 1)

    def report(tag, out):
        roi, xb, yb, xid, yid = out
        e0 = float((xb[..., :3] - rxb[..., :3]).abs().max() / rxb[..., :3].abs().max())
        r0 = float((xb[..., :3] - rxb[..., :3]).pow(2).mean().sqrt() / rxb[..., :3].pow(2).mean().sqrt())
        agree = float(((xid == ref[4]) & (yid == ref[5])).float().mean())
        agree0 = float((((xid >> 3) == (ref[4] >> 3)) & ((yid >> 3) == (ref[5] >> 3))).float().mean())
        print(f"{tag:34s} init x-logits max/max {e0:.2e} rms/rms {r0:.2e} | init-cell agree {agree0:.4f} | 64x64 cell agree {agree:.4f}")

    report("fp32 factored", head(feats, sd, idx, N, None))
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        report(f"{name} all", head(feats, sd, idx, N, dt))
        report(f"{name} weights exact", head(feats, sd, idx, N, dt, weights_exact=True))
        report(f"{name} image branch exact", head(feats, sd, idx, N, dt, img_exact=True))
        report(f"{name} z exact", head(feats, sd, idx, N, dt, z_exact=True))
        report(f"{name} img+w exact", head(feats, sd, idx, N, dt, img_exact=True, weights_exact=True))


if __name__ == "__main__":
    main()

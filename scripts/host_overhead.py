"""Host enqueue time vs GPU time of one head step (is the step launch-bound?)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from checkerpose_b200 import head, synthetic as syn
from checkerpose_b200.model import init, pipeline
from checkerpose_b200.model.backbone import FeatureListBackbone

torch.set_grad_enabled(False)
if os.environ.get("CUDNN_BENCHMARK"):
    torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
B, N = int(os.environ.get("KB_B", 256)), 4096
head.set_compute_dtype(torch.bfloat16)
g = torch.Generator().manual_seed(3)
p3d = syn.p3d_normed_tensor(syn.load_fps_xyz("lmo", 1, N)).to(dev)
sd = syn.synthetic_state_dict(syn.head_param_spec(N), g)
inet = init.InitNet_GNN(npoint=N, p3d_normed=p3d, res_log2=3, backbone_name="hrnet_w18", pretrain_backbone=False,
                        max_batch_size=B, num_graph_module=2, graph_k=20, img_backbone=FeatureListBackbone())
net = pipeline.PoseNet_GNNskip(inet, npoint=N, p3d_normed=p3d, res_log2=6, num_filters=256, max_batch_size=B, local_k=2,
                               leaky_slope=0.01, num_graph_module=3, graph_k=20)
net.load_state_dict(sd, strict=True)
net = net.to(dev).eval()
gg = torch.Generator(device=dev).manual_seed(1)
feats = [torch.relu(torch.randn(B, c, s, s, generator=gg, device=dev)).to(torch.bfloat16) for c, s in zip(syn.HRNET_W18_DIMS, syn.HRNET_W18_SIZES)]
bbox = syn.synthetic_bboxes(B, torch.Generator().manual_seed(9)).to(dev)
pexp = p3d.expand(B, -1, -1)
for _ in range(3):
    net.forward_with_correspondences(feats, pexp, bbox)
torch.cuda.synchronize()
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(K):
    net.forward_with_correspondences(feats, pexp, bbox)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: host enqueue {1e3 * (t1 - t0) / K:.2f} ms/step, GPU {e0.elapsed_time(e1) / K:.2f} ms/step, wall {1e3 * (t2 - t0) / K:.2f} ms/step")

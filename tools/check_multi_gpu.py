"""torchrun --nproc-per-node 2 tools/check_multi_gpu.py: the sharded detector (all-gather + NMS sharded over the
gathered set + all-gather of the kept rows) must reproduce, on every rank, what one engine computes for the whole
global batch (development aid / multi-GPU validation)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from m3dssd_b200 import synth
from m3dssd_b200.model.M3d_inference_align import build
from m3dssd_b200.parallel import ShardedDetector, shard_range

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
LB, H, W = 2, 96, 320
conf = synth.make_conf(attention=None, center_align=True, shape_align=True, crop_size=(H, W))
net = build(conf, "test")
synth.randomize_weights(net)
net = net.cuda().eval()
images = synth.make_images(world * LB, (H, W)).cuda()
lo, hi = shard_range(world * LB, world, rank)
det = ShardedDetector(net, LB, H, W, precision="bf16", use_graph=False)
kept, num = det.step(images[lo:hi])
torch.cuda.synchronize()
kept, num = kept.clone(), num.clone()
dets_all, det_num_all = det.dets_all.clone(), det.det_num_all.clone()  # the gathered pre-NMS set of config 5
ref = net.engine(LB, H, W, precision="bf16", use_graph=False)
ok = True
for r in range(world):
    a, b = shard_range(world * LB, world, r)
    k, n = ref.detect(images[a:b])
    torch.cuda.synchronize()
    ok &= bool(torch.equal(n, num[a:b])) and bool(torch.equal(k, kept[a:b]))
    ok &= bool(torch.equal(ref.dets, dets_all[a:b])) and bool(torch.equal(ref.det_num, det_num_all[a:b]))
print("rank %d: sharded result == single-engine result on all %d images: %s (kept per image %s)" % (rank, world * LB, ok, num.tolist()), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

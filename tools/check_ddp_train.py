"""torchrun --nproc-per-node 2 tools/check_ddp_train.py: two ranks train the native path (C-ABI convolutions / DCNv2 in
both directions) on different image shards under DistributedDataParallel; after each step the parameters must be
identical on both ranks (the all-reduced gradients were), and the loss must go down."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from m3dssd_b200 import synth, train
from m3dssd_b200.model.M3d_inference_align import build

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr_ = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
conf = synth.make_conf(attention=None, center_align=False, shape_align=False, crop_size=(96, 320), batch_size=2)
net = build(conf, "train")
synth.randomize_weights(net)
net = net.cuda()
train.enable(net)
model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[lr_], find_unused_parameters=True)
step = train.TrainStep(model, conf, lr=0.002, native=False)
x = synth.make_images(2, (96, 320), seed=10 + rank).cuda()
labels, t2, t3 = train.surrogate_targets(conf, 2, "cuda", seed=rank, fg_per_image=60)
losses = [float(step(x, labels, t2, t3).detach()) for _ in range(6)]
flat = torch.cat([p.detach().flatten() for p in net.parameters()])
ref = flat.clone()
dist.broadcast(ref, src=0)
same = bool(torch.equal(flat, ref))
ok = same and losses[-1] < losses[0]
print("rank %d: parameters identical across ranks after 6 DDP steps: %s; loss %.4f -> %.4f" % (rank, same, losses[0], losses[-1]), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

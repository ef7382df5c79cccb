"""Manual multi-GPU check (not collected by pytest; needs >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/multi_gpu_allreduce_check.py

The statistics all-reduce of the path runs through the C ABI's own NCCL communicator (fad_comm_unique_id / fad_comm_init /
fad_stats_allreduce, include/fadtk_b200.h).  Every rank accumulates its shard of one fp16 matrix with the shared shift;
after the native all-reduce the finalised (mu, cov) must equal the single-process statistics of the whole matrix and be
bit-identical to what torch.distributed.all_reduce gives for the same accumulators."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from fadtk_b200 import _native, dist          # noqa: E402
from fadtk_b200.utils import DeviceStatistics  # noqa: E402


def main():
    dist.init_from_env("nccl")
    r, w = dist.rank(), dist.world_size()
    eng = _native.engine(int(os.environ.get("LOCAL_RANK", "0")))
    dev = eng.torch_device
    rng = np.random.default_rng(7)
    for d, n in ((128, 20011), (768, 5003)):
        rows = (rng.standard_normal((n, d)) * rng.uniform(0.3, 2.0, d) + rng.standard_normal(d)).astype(np.float16)
        mine = dist.shard(list(range(n)), r, w)
        part = torch.from_numpy(rows[mine[0]:mine[-1] + 1]).to(dev)
        shift = torch.from_numpy(rows[:64].astype(np.float32).mean(0).astype(np.float16)).to(dev)
        acc = eng.stats_accumulate(part, shift, eng.stats_new(d))
        via_torch = acc.clone()
        torch.distributed.all_reduce(via_torch)
        assert dist.enable_native_allreduce(eng) and eng.has_comm
        dist.allreduce_sum_(acc)                              # ncclAllReduce issued by libfadtk_b200.so
        torch.cuda.synchronize()
        assert torch.equal(acc, via_torch), (acc - via_torch).abs().max().item()
        mu, cov = eng.stats_finalize(acc, shift, d)
        x = rows.astype(np.float64)
        assert acc[0].item() == n
        assert np.abs(mu.cpu().numpy() - x.mean(0)).max() < 1e-10
        ref = np.cov(x, rowvar=False)
        assert np.abs(cov.cpu().numpy() - ref).max() < 1e-11 * np.abs(ref).max()
        # the host-facing wrapper takes the same route
        st = DeviceStatistics(d, eng, reduce_ranks=True)     # rank 0's shift is broadcast on the first add
        st.add(part)
        st.allreduce()
        assert st.count() == n
        mu2, cov2 = st.finalize()
        assert np.abs(mu2.cpu().numpy() - x.mean(0)).max() < 1e-10
        assert np.abs(cov2.cpu().numpy() - ref).max() < 1e-11 * np.abs(ref).max()
    sys.stdout.write(f"[rank{r}: native all-reduce ok over {w} GPUs]\n")
    sys.stdout.flush()
    dist.shutdown()


if __name__ == "__main__":
    main()

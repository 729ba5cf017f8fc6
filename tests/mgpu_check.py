"""Run under torchrun (any world >= 2): sharded DLRM == single-GPU DLRM on the same global batch
(predictions, loss, every table shard, Adam moments and MLP weight after 3 training steps with
eval forwards interleaved).  The logic lives in sharded.parity_self_check — the same function
bench.py runs untimed before a multi-GPU measurement."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recommend_tf2_b200 as pkg  # noqa: E402,F401
from recommend_tf2_b200.sharded import parity_self_check  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    torch.backends.cuda.matmul.allow_tf32 = False
    mode = os.environ.get("RTF_EXCHANGE", "peer")
    res = parity_self_check(mode, os.environ.get("RTF_PEER_GATHER", "owner"),
                            int(os.environ.get("RTF_REPLICATE", "70")))
    if rank == 0:
        print(("mgpu_check ok: " if res["ok"] else "mgpu_check FAILED: ") + json.dumps(res), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()

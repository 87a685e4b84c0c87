#!/usr/bin/env python
"""The multi-rank paths on real GPUs over NCCL (the CPU suite covers the same code with gloo + host emulation):
sharded spectrum sweep with an all-gather of (R, T), and Brillouin-zone-integrated field maps with k-points sharded over the
ranks and ONE all-reduce of the summed maps.  Checked against the goldens of the unmodified reference.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/nccl_bzi_check.py"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import cases  # noqa: E402
from tests.util import build_crystal, engine, gold  # noqa: E402
from khepri_b200.sharding import allreduce_sum, sweep_sharded  # noqa: E402
from khepri_b200.beams import bzi_fields  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = engine("cuda")
st, srcs = cases.case_suh03()
cl = build_crystal(st, eng)
wl = np.array([s["wavelength"] for s in srcs])
R, T = sweep_sharded(cl, wl, te=1.0, tm=0.0)
err_rt = float(np.abs(np.stack([R, T], 1) - gold("suh03")["RT"]).max())
tot = allreduce_sum(torch.full((3,), complex(rank + 1, -rank), dtype=torch.complex128, device="cuda"))
stb, c = cases.case_bzi_beam()
clb = build_crystal(stb, eng, fields=True)
xo, yo, zo = c["out"]
E, H = bzi_fields(clb, c["wl"], c["kbz"], gold("bzi_beam")["amplitudes"].reshape(len(c["kbz"]), -1), xo, yo, zo)
gb = gold("bzi_beam")["fields"]
err_f = float(np.abs(np.asarray((E, H)) - gb).max() / np.abs(gb).max())
ok = err_rt <= 1e-9 and err_f <= 1e-9 and np.allclose(tot.cpu().numpy(), sum(complex(r + 1, -r) for r in range(world)))
dist.barrier()
if rank == 0:
    print(json.dumps({"check": "NCCL sharded sweep + BZI field sum", "world": world, "backend": dist.get_backend(), "sources": int(wl.size),
                      "max_abs_err_RT_vs_reference_golden": err_rt, "rel_err_fields_vs_reference_golden": err_f, "ok": bool(ok)}))
dist.destroy_process_group()
sys.exit(0 if ok else 1)

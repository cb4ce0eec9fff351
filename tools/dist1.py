"""World-size-1 run of the multi-GPU code path (sample -> select/partition -> exchange-with-self -> local pipeline),
so that its kernels can be profiled with ncu on one GPU.  python tools/dist1.py [steps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1"); os.environ.setdefault("LOCAL_RANK", "0")
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
import smj_b200
from smj_b200 import smj as S

rank, world, local = smj_b200.dist.init()
n = 10_000_000
cfg = S.default_config(select_val1=3 * n // 2, select_val2=3 * n // 2, nr_gpus=1)
d1, d2 = smj_b200.synth_device_table(n, 4, 1), smj_b200.synth_device_table(n, 4, 2)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    out, st = smj_b200.run(d1, d2, cfg=cfg, on_device=True, keep_output=True)
    smj_b200.lib().smj_table_free(C.byref(out))
print({k: round(v, 3) if isinstance(v, float) else v for k, v in st.items() if k.endswith("_ms") or k.startswith("rows")})
smj_b200.lib().smj_shutdown()

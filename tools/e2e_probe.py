import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pyarrow as pa
from fugue_b200 import api as fa
from fugue_b200.schema import Schema
from fugue_b200.table import B200Table
from fugue_b200.dataframe import ArrowDataFrame, B200DataFrame
from fugue_b200.partition import PartitionSpec
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
SCHEMA = "key:long,i1:long,i2:long,i3:long,v0:double,v1:double,v2:double,v3:double"
cols = [torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device=dev)] + [torch.randint(-(2**62), 2**62, (n,), dtype=torch.int64, device=dev) for _ in range(3)] + [torch.randn(n, dtype=torch.float64, device=dev) for _ in range(4)]
hc = []
for c in cols:
    h = torch.empty(c.shape, dtype=c.dtype, pin_memory=True); h.copy_(c); hc.append(h)
torch.cuda.synchronize()
del cols
sch = Schema(SCHEMA)
host_table = pa.Table.from_arrays([pa.Array.from_buffers(tp, n, [None, pa.py_buffer(h.numpy())]) for h, tp in zip(hc, sch.types)], names=sch.names)
e = fa.make_execution_engine("b200")
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); d = e.to_df(ArrowDataFrame(host_table)); t1 = T()
    r = e.repartition(d, PartitionSpec(by="key", num=256)); t2 = T()
    a = r.as_arrow(); t3 = T()
    print(f"iter {it}: to_df {1e3*(t1-t0):.1f} ms ({6.4/(t1-t0):.1f} GB/s)  partition {1e3*(t2-t1):.1f} ms  as_arrow {1e3*(t3-t2):.1f} ms ({6.4/(t3-t2):.1f} GB/s)")
    del d, r, a
# raw copies
x = torch.empty(n, dtype=torch.int64, device=dev)
t0 = T(); x.copy_(hc[0], non_blocking=True); t1 = T(); print(f"raw H2D 0.8GB pinned: {1e3*(t1-t0):.1f} ms {0.8/(t1-t0):.1f} GB/s")
import numpy as np
v = torch.from_numpy(np.frombuffer(host_table.column(0).chunk(0).buffers()[1], dtype='i8'))
print("is_pinned via arrow view:", v.is_pinned())
t0 = T(); x.copy_(v, non_blocking=True); t1 = T(); print(f"view H2D: {1e3*(t1-t0):.1f} ms")
h = torch.empty(n, dtype=torch.int64, pin_memory=True)
t0 = T(); h.copy_(x, non_blocking=True); t1 = T(); print(f"raw D2H: {1e3*(t1-t0):.1f} ms {0.8/(t1-t0):.1f} GB/s")
t0 = T(); h2 = torch.empty(n, dtype=torch.int64, pin_memory=True); t1 = T(); print(f"pinned alloc 0.8GB: {1e3*(t1-t0):.1f} ms")
del h2
t0 = T(); h2 = torch.empty(n, dtype=torch.int64, pin_memory=True); t1 = T(); print(f"pinned alloc again: {1e3*(t1-t0):.1f} ms")

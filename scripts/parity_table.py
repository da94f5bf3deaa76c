"""Writes, for the full-size bf16 step, every quantity's error (max-norm relative, relative L2, gain) of
  product vs exact fp64 oracle | product vs bf16-storage oracle | bf16-storage oracle vs exact oracle
and runs the assertions of tests/test_parity_bf16_gpu.py::check_bf16_step.
usage: python scripts/parity_table.py [batch] [graph:0|1] [out.txt]"""
import sys
sys.path.insert(0, ".")
from tests.test_parity_bf16_gpu import check_bf16_step

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
graph = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/r2_parity_table.txt"
check_bf16_step(B, graph, table=out)
print(open(out).read())

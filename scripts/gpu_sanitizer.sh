#!/bin/bash
# compute-sanitizer over the kernels with inter-CTA protocols and aliased shared memory:
#   memcheck + racecheck of (a) the toy order-2 lift + DBGNN forward of __graft_entry__.smoke(), (b) a chain build with the
#   heavy-row fallback forced on, (c) a multi-tile onesweep sort, (d) the staged tcgen05 GCN layer (CASES="gcn" runs only
#   the named cases).  Run under gpurun on one B200; logs go to gpurun_out/.
#     gpurun --timeout 1500 -- 'bash scripts/gpu_sanitizer.sh r02'
set -u
tag=${1:-san}
out=gpurun_out
mkdir -p $out
cat > /tmp/san_case.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import pathpyg_b200 as pp
from pathpyg_b200 import ops
import __graft_entry__ as g
which = sys.argv[1]
dev = torch.device("cuda", 0)
if which == "smoke":
    g.smoke()
elif which == "chain":
    gen = torch.Generator().manual_seed(1)
    n, m = 300, 20000
    ei = torch.randint(0, n, (2, m), generator=gen).to(dev)
    t = torch.sort(torch.randint(0, 400, (m,), generator=gen)).values.to(dev)
    for heavy in ("256", "3"):
        os.environ["PPG_CHAIN_HEAVY"] = heavy
        model = pp.MultiOrderModel.from_temporal_graph(pp.TemporalGraph.from_tensors(ei, t, n), delta=4, max_order=4)
        print({k: (v.n, v.m) for k, v in model.layers.items()})
elif which == "dist":
    import torch.distributed as dist
    from pathpyg_b200 import parallel
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29571")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    gen = torch.Generator().manual_seed(3)
    n, m = 200, 12000
    ei = torch.randint(0, n, (2, m), generator=gen).to(dev)
    t = torch.sort(torch.randint(0, 300, (m,), generator=gen)).values.to(dev)
    for heavy in ("256", "3"):
        os.environ["PPG_CHAIN_HEAVY"] = heavy
        layers = parallel.distributed_temporal_layers(ei, t, n, 4, 4)
        print({k: (v.num_nodes, v.edge_index.size(1)) for k, v in layers.items()})
    dist.destroy_process_group()
elif which == "gcn":
    # the staged tcgen05 GCN layer (producer / consumer / MMA warps, mbarrier hand-overs, cp.async stages): several
    # tiles per CTA, bit-identical to the single-role kernel
    from pathpyg_b200 import _lib
    gen = torch.Generator().manual_seed(4)
    os.environ["PPG_GCN_TC_GRID"] = "6"          # 6 CTAs x 4-5 tiles
    n, e, F, H = 128 * 26 + 50, 9000, 32, 32
    ei = torch.randint(0, n, (2, e), generator=gen)
    ei[1, :600] = torch.randint(0, 4, (600,), generator=gen)
    graph = ops.gcn_prepare(ei.to(dev), (torch.rand(e, generator=gen) + 0.5).to(dev), n)
    x, W, b = torch.randn(n, F, generator=gen).to(dev), (torch.randn(H, F, generator=gen) / 6).to(dev), torch.randn(H, generator=gen).to(dev)
    outs = []
    for v in ("staged", "single"):
        os.environ["PPG_GCN_TC"] = v
        outs.append(ops.gcn_layer_tc(graph, x, W, b, _lib.ACT_ELU))
    assert torch.equal(outs[0], outs[1])
elif which == "sort":
    gen = torch.Generator().manual_seed(2)
    keys = torch.randint(0, 1 << 40, (20000,), generator=gen).to(dev)
    want = torch.sort(keys, stable=True)
    perm, _ = ops.sort_pairs_u64(keys, 40)
    assert torch.equal(keys, want.values) and torch.equal(perm.long(), want.indices)
torch.cuda.synchronize()
print(which, "ok")
PY
for tool in memcheck racecheck; do
  for c in ${CASES:-smoke chain dist sort gcn}; do
    timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py $c > $out/${tag}_sanitizer_${tool}_${c}.log 2>&1
    echo "$tool $c: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_sanitizer_${tool}_${c}.log | tail -1)"
  done
done

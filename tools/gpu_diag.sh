# diagnostics: per-chain cycle counts of block 0 in a full launch (variant built with -DSQ_DEBUG_TIMING), PCIe copy rates
mkdir -p gpurun_out/diag
keep=/tmp/keep_lib.so; cp fastqueeze_b200/libseqarc_b200.so $keep
cp _variants/ticks.so fastqueeze_b200/libseqarc_b200.so
for nb in 274 74 1; do
  echo "== blocks $nb" >> gpurun_out/diag/ticks.log
  timeout 400 python bench.py --blocks $nb --steps 1 --warmup 1 --no-cpu --no-e2e 2>/dev/null | grep "dec warp" | tail -4 >> gpurun_out/diag/ticks.log
done
cp $keep fastqueeze_b200/libseqarc_b200.so
python - > gpurun_out/diag/pcie.log 2>&1 <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, f in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * n / 1e9 / (time.perf_counter() - t)))
h2 = torch.empty(50 << 20, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(50 << 20, dtype=torch.uint8, device="cuda")
d2.copy_(h2, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(40): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
print("h2d 50 MiB pieces %.1f GB/s" % (40 * (50 << 20) / 1e9 / (time.perf_counter() - t)))
PY
cat gpurun_out/diag/ticks.log gpurun_out/diag/pcie.log

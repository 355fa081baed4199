#!/bin/bash
# run_variants.sh — build kernel variants here, measure them back to back on ONE GPU box, print a table.
#
#   profiles/run_variants.sh OUT_PREFIX "name1:-DFLAG=1 -DOTHER=2" "name2:-DFLAG=3" ...
#
# Step 1 (this container, no GPU): every variant is cross-compiled into griddy_b200/csrc/variants/ with
#   griddy_b200/csrc/build_variant.sh; "main" (the in-tree library, no flags) is always measured first.
# Step 2 (one gpurun call): for each variant  GRIDDY_CUDA_LIB=… python bench.py --no-cpu-baseline --e2e-steps 5,
#   bench lines to gpurun_out/OUT_PREFIX_<name>.json.
# Step 3: a markdown table of ms per tick and per-kernel ms (events around every launch) on stdout, ready for
#   profiles/r*_variants.md.  Add  TESTS=1  to run `pytest -m gpu -x` on the in-tree library first.
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
prefix="$1"; shift
names="main"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  "$root/griddy_b200/csrc/build_variant.sh" "$name" $flags >/dev/null
  names="$names $name"
done
cat > "$root/.run_variants.sh" <<EOS
[ -n "$TESTS" ] && { python -m pytest tests -m gpu -x -q 2>&1 | tail -2; }
for v in $names; do
  if [ \$v = main ]; then unset GRIDDY_CUDA_LIB; else export GRIDDY_CUDA_LIB=\$PWD/griddy_b200/csrc/variants/libgriddy_cuda_\$v.so; fi
  python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 5 > gpurun_out/${prefix}_\$v.json 2> gpurun_out/${prefix}_\$v.err || echo "\$v failed"
done
EOS
(cd "$root" && gpurun --timeout 900 -- 'bash .run_variants.sh' | tail -3)
rm -f "$root/.run_variants.sh"
python - "$root" "$prefix" $names <<'PY'
import json, sys
root, prefix, names = sys.argv[1], sys.argv[2], sys.argv[3:]
print("| variant | tick | k_advance | k_slow | k_unlink | k_link | e2e G/s |\n|---|---|---|---|---|---|---|")
for v in names:
    try:
        d = json.loads(open(f"{root}/gpurun_out/{prefix}_{v}.json").read().strip().splitlines()[-1])
        k = d["roofline"]["kernel_ms_per_tick"]
        print(f"| {v} | {d['ms_per_step']:.4f} | {k['k_advance']:.4f} | {k['k_slow']:.4f} | {k['k_unlink']:.4f} | {k['k_link']:.4f} | {d['e2e']['value'] / 1e9:.2f} |")
    except Exception as e:  # noqa: BLE001
        print(f"| {v} | failed: {e!r} |")
PY

python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
seen=False
for l in sys.stdin.read().strip().splitlines():
    if l.startswith('[xnb]') and not seen: print(l); seen=True
    if l.startswith('{'):
        d=json.loads(l); print({k:round(v,4) for k,v in d['breakdown_ms_per_step'].items()}, round(d['ms_per_step'],4))
"; }
for v in "XNB_TILE_DEBUG=1" "XNB_TILE_I=3 XNB_TILE_J=2" "XNB_TILE_I=4 XNB_TILE_J=2" "XNB_TILE_I=2 XNB_TILE_J=2" "XNB_TILE_I=4 XNB_TILE_J=1" "XNB_TILE_I=4 XNB_TILE_J=3"; do
  echo "== $v"; run XNB_TILE_DEBUG=1 $v
done
ncu --set full --clock-control none --import-source on -k regex:"k_lj_force_tiled" -s 3 -c 1 -f -o gpurun_out/prof_r1e python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu5.log 2>&1

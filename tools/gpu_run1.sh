for k in k_fast k_quadtree k_blur k_resize; do
ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip 8 --launch-count 1 -o gpurun_out/r1b_$k -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/c_$k.log 2>&1
done
for L in 1 2 3 4 8; do echo "LANES=$L"; HYORB_LANES=$L python bench.py --steps 20 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:round(v['ms_per_step'],3) for k,v in d['stages'].items()})"; done
echo SIDE; HYORB_SIDE_BLUR=1 python bench.py --steps 20 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"
for P in 8 16 64 128; do echo "PAIRS=$P"; python bench.py --steps 20 --warmup 3 --no-cpu --pairs $P | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"; done

#!/bin/bash
# Turns the artefacts of one profile run (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
#   bash tools/refresh_profiles.sh r2f
set -e
tag=${1:-r2f}
cp gpurun_out/${tag}_bench.json profiles/${tag}_bench.json
cp gpurun_out/${tag}_bench_reference.json profiles/
cp gpurun_out/${tag}_launches.csv profiles/
( python tools/ncu_summary.py launches profiles/${tag}_launches.csv; echo
  python tools/ncu_summary.py launches profiles/${tag}_launches.csv 0 84 "[launches 0-83 of the library = warm-up + timed device-resident steps of bench.py: 6 steps x 14 kernels over 256 images, 1 lane -- the launches the bench's stage times describe]" ) > profiles/${tag}_launches.txt
for k in k_fast k_level k_describe k_quadtree k_stereo_search; do
  python tools/ncu_summary.py full gpurun_out/${tag}_$k.ncu-rep > profiles/${tag}_${k}_full.txt
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for v in rows[2:]:
    out=[]
    for i,k in enumerate(h):
        if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k:
            try: x=float(v[i].replace(',',''))
            except: continue
            if x>0: out.append((x,k))
    tot=sum(x for x,_ in out)
    print('# warp stall samples (smsp__pcsamp_warps_issue_stalled_*), share of all samples')
    for x,k in sorted(out,reverse=True)[:8]: print('%-28s %6d  %.3f'%(k.replace('smsp__pcsamp_warps_issue_stalled_',''),x,x/tot))
" >> profiles/${tag}_${k}_full.txt
done
python - "$tag" <<'PY'
import subprocess,csv,json,sys
tag=sys.argv[1]
IM=256
out={"_what":"DRAM traffic per kernel from `ncu --set full` captures of bench.py's own step (128 pairs = 256 images per launch, 1 lane): dram__bytes_read.sum + dram__bytes_write.sum; k_level = sum over the eight per-level launches of one step"}
tot=0
for k in ["k_fast","k_level","k_describe","k_quadtree","k_stereo_search"]:
    o=subprocess.run(["ncu","-i",f"gpurun_out/{tag}_{k}.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(o.splitlines())); h,u=rows[0],rows[1]
    def val(v,name):
        i=h.index(name); x=float(v[i].replace(',','')); unit=u[i]
        return x*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[unit]
    rd=sum(val(v,'dram__bytes_read.sum') for v in rows[2:]); wr=sum(val(v,'dram__bytes_write.sum') for v in rows[2:])
    out[k]={"dram_bytes_per_image":(rd+wr)/IM,"dram_bytes_read":rd,"dram_bytes_write":wr,"images_in_capture":IM,"launches_in_capture":len(rows)-2,
            "source":f"profiles/{tag}_{k}_full.txt"}
    tot+=(rd+wr)/IM
out["_sum_dram_bytes_per_image"]=tot
json.dump(out,open("profiles/traffic.json","w"),indent=1)
print("DRAM bytes per image, all kernels:", round(tot))
PY
python tools/sass_summary.py > profiles/${tag}_sass_summary.txt

#!/bin/bash
# Turns the artefacts of one profile run (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
#   bash tools/refresh_profiles.sh r1e
set -e
tag=${1:-r1e}
cp gpurun_out/${tag}_bench.json profiles/${tag}_bench.json
cp gpurun_out/${tag}_bench_reference.json profiles/
cp gpurun_out/${tag}_launches.csv profiles/
( python tools/ncu_summary.py launches profiles/${tag}_launches.csv; echo
  python tools/ncu_summary.py launches profiles/${tag}_launches.csv 168 224 "[launches 168-223 of the library = bench.py's serialised pass: 4 steps x 14 kernels over 256 images]"; echo
  python tools/ncu_summary.py launches profiles/${tag}_launches.csv 0 168 "[launches 0-167 = warm-up + timed device-resident steps: 6 steps x 2 lanes x 14 kernels over 128 images]" ) > profiles/${tag}_launches.txt
for k in k_fast k_resize k_blur k_describe k_quadtree k_stereo_search; do
  python tools/ncu_summary.py full gpurun_out/${tag}_$k.ncu-rep > profiles/${tag}_${k}_full.txt
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
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
out={}
for k in ["k_fast","k_resize","k_blur","k_describe","k_quadtree","k_stereo_search"]:
    o=subprocess.run(["ncu","-i",f"gpurun_out/{tag}_{k}.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(o.splitlines())); h,u,v=rows[0],rows[1],rows[2]
    def val(name):
        i=h.index(name); x=float(v[i].replace(',','')); unit=u[i]
        return x*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[unit]
    rd,wr=val('dram__bytes_read.sum'),val('dram__bytes_write.sum')
    out[k]={"dram_bytes_per_image":(rd+wr)/64,"dram_bytes_read":rd,"dram_bytes_write":wr,"images_in_capture":64,
            "source":f"profiles/{tag}_{k}_full.txt (ncu --set full, one launch of 64 C2 images: bench.py --pairs 32, 1 lane)"}
out["k_resize"]["note"]="one of the 7 per-level launches (launch-skip 4), not the whole pyramid"
json.dump(out,open("profiles/traffic.json","w"),indent=1)
PY

#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"
  DATOK_B200_LIB=$PWD/build_variants/$v.so python scripts/profile_one.py $((1<<30)) de 2>&1 | tail -2 | head -1
  DATOK_B200_LIB=$PWD/build_variants/$v.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_write_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_write_lookup_hit.sum --clock-control none -k regex:"walk_fused|stitch|compact" --csv --log-file gpurun_out/tr_$v.csv python scripts/profile_one.py $((1<<30)) de > gpurun_out/tr_$v.log 2>&1
  python - <<P
import csv
rows=[r for r in csv.reader(open('gpurun_out/tr_$v.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]),r[ki].split('(')[0][:34]),{})[r[mi]]=r[vi]
for k in sorted(d)[-9:]:
    print(k, {m.replace('lts__t_sectors_srcunit_tex_op_write_lookup_','w').replace('.sum',''):v for m,v in d[k].items()})
P
done

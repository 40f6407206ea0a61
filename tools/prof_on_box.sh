#!/bin/bash
# usage: tools/prof_on_box.sh <tag> <kernel-regex> <prof_kernels args...>
# ncu --set full on the box; the report stays in /tmp unless small, the text summary goes to gpurun_out/.
tag=$1; regex=$2; shift 2
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$regex" -o /tmp/$tag -f python tools/prof_kernels.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/$tag.ncu-rep > gpurun_out/${tag}_summary.txt 2>&1
sz=$(stat -c%s /tmp/$tag.ncu-rep)
echo "report bytes: $sz"
if [ "$sz" -lt 30000000 ]; then cp /tmp/$tag.ncu-rep gpurun_out/; fi
tail -2 gpurun_out/${tag}_ncu.log

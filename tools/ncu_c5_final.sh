NCU="ncu --set full --clock-control none -f"
mkdir -p gpurun_out
name=r2_c5_filter_final
$NCU -k regex:"prepare|atrous|taaKernel|presentKernel" --launch-skip 16 -c 8 -o gpurun_out/$name python tools/profile_target.py c5 3 > gpurun_out/ncu_$name.log 2>&1
{ echo "# final build (one instantiation of the history-reading kernels, per-row ownership test): ncu --set full --clock-control none -k regex:\"prepare|atrous|taaKernel|presentKernel\" --launch-skip 16 -c 8 python tools/profile_target.py c5 3"; python tools/ncu_summary.py gpurun_out/$name.ncu-rep; } > gpurun_out/$name.txt
rm -f gpurun_out/$name.ncu-rep
wc -l gpurun_out/$name.txt

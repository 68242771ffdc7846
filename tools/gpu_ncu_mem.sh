# memory behaviour of k_dec_main in the full 274-block launch, per library variant under _variants/
mkdir -p gpurun_out/mem
keep=/tmp/keep_lib.so; cp fastqueeze_b200/libseqarc_b200.so $keep
CMD="python bench.py --blocks ${1:-274} --steps 1 --warmup 1 --no-cpu --no-e2e"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
M=$M,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum
M=$M,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st_lookup_hit.sum
M=$M,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_write.sum
M=$M,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum,l1tex__m_xbar2l1tex_read_sectors.sum
for v in _variants/*.so; do
  n=$(basename $v .so)
  cp $v fastqueeze_b200/libseqarc_b200.so
  timeout 900 ncu --metrics $M --clock-control none -k regex:k_dec_main --launch-skip 1 -c 1 --csv --log-file gpurun_out/mem/$n.csv $CMD > gpurun_out/mem/$n.log 2>&1
  grep k_dec_main gpurun_out/mem/$n.csv | awk -F'","' '{print $13, $15}' | tr -d '"' | sed "s/^/$n /"
done
cp $keep fastqueeze_b200/libseqarc_b200.so

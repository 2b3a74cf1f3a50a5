for n in t0 t1; do
  export GO2P_LIB=$PWD/go2_onnx_controller_b200/lib/exp_$n.so
  python scripts/wide_one.py 75776 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/wide_layers_$n.csv python scripts/wide_one.py 75776 > /dev/null 2>&1
  echo "$n exit $?"
done

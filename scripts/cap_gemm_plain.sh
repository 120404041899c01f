# ncu --set full of the CTA-pair persistent GEMM at the decode QKV shape and the training FFN1 (plain) shape; text pages only
O=gpurun_out; mkdir -p $O
export_rep() { ncu -i $1.ncu-rep --page details > $1_ncu_details.txt 2>&1; ncu -i $1.ncu-rep --page raw --csv > $1_ncu_raw.csv 2>&1; rm -f $1.ncu-rep; }
python scripts/one_gemm.py 30000 1536 512 256 > $O/plain_q.log 2>&1; cat $O/plain_q.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist -s 3 -c 1 -o $O/r02_gemm_pair_decode_qkv_v10 -f python scripts/one_gemm.py 30000 1536 512 256 > $O/ncu_q.log 2>&1
echo "exit $?"; export_rep $O/r02_gemm_pair_decode_qkv_v10
python scripts/one_gemm.py 41472 2048 512 256 > $O/plain_f.log 2>&1; cat $O/plain_f.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_persist -s 3 -c 1 -o $O/r02_gemm_pair_train_ffn1_plain_v10 -f python scripts/one_gemm.py 41472 2048 512 256 > $O/ncu_f.log 2>&1
echo "exit $?"; export_rep $O/r02_gemm_pair_train_ffn1_plain_v10
grep -h "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed\|gpu__time_duration.sum\|lts__t_bytes.sum \|dram__bytes_read.sum \|l1tex__m_xbar2l1tex_read_bytes.sum " $O/r02_gemm_pair_*_v10_ncu_raw.csv | head

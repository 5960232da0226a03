O=gpurun_out
echo "== 4K 70 regs"; python tools/sweep_plan.py --width 3840 --height 2160 --steps 6 "0:256:3:16,0:256:1:0" "0:256:3:16,0:256:2:0" "0:256:3:12,0:256:2:0" "0:256:3:20,0:256:2:0" 2>&1 | tee $O/sweep_plan_r2h_4k_70.txt
echo "== 4K 40 regs"; RMB_MARCH_MIN_BLOCKS=6 python tools/sweep_plan.py --width 3840 --height 2160 --steps 6 "0:256:2:16,0:256:1:0" "0:256:4:16,0:256:2:0" "0:256:6:16,0:256:2:0" "0:256:6:16,0:256:3:0" "0:256:6:0" "64:256:6:16,0:256:2:0" 2>&1 | tee $O/sweep_plan_r2h_4k_40.txt
echo "== 4K 48 regs (min blocks 5)"; RMB_MARCH_MIN_BLOCKS=5 python tools/sweep_plan.py --width 3840 --height 2160 --steps 6 "0:256:5:16,0:256:2:0" "0:256:4:16,0:256:2:0" 2>&1 | tee $O/sweep_plan_r2h_4k_48.txt
echo "== 4K min blocks 4"; RMB_MARCH_MIN_BLOCKS=4 python tools/sweep_plan.py --width 3840 --height 2160 --steps 6 "0:256:4:16,0:256:2:0" "0:256:3:16,0:256:2:0" 2>&1 | tee $O/sweep_plan_r2h_4k_mb4.txt
echo "== 1080 min blocks 4"; RMB_MARCH_MIN_BLOCKS=4 python tools/sweep_plan.py "0:256:2:16,0:256:1:0" "0:256:3:16,0:256:1:0" "0:256:4:16,0:256:2:0" 2>&1 | tee $O/sweep_plan_r2h_1080_mb4.txt
echo "== 1080 70 regs more"; python tools/sweep_plan.py "0:256:2:16,0:256:1:0" "0:256:2:12,0:256:1:0" "0:256:2:20,0:256:1:0" "0:256:2:24,0:256:1:0" "0:256:3:16,0:256:1:0" "0:256:2:16,0:256:1:8,0:128:1:0" "0:256:2:16,0:128:2:0" 2>&1 | tee $O/sweep_plan_r2h_1080_70.txt

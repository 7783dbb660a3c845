"""profiles/r2_ncu_bench_scale.json from one `ncu --set full` capture of bench.py's own launch sequence:
    ncu --set full --clock-control none -k regex:'static_fast_kernel|static_solve_kernel|static_curves_pos_kernel|chain_eval_group_kernel' \
        -s 4 -c 4 -o X python bench.py --steps 1 --warmup 1 --no-extra --no-cpu-baseline
usage: ncu_bench_scale.py X.ncu-rep <workload> <configs per step> > profiles/r2_ncu_bench_scale.json"""
import json
import subprocess
import sys

rep, workload, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
summ = json.loads(subprocess.run([sys.executable, __file__.replace("ncu_bench_scale.py", "ncu_summary.py"), rep],
                                 capture_output=True, text=True).stdout)


def num(s):
    v, unit = s.split()[0], (s.split() + [""])[1].lower()
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(unit, 1)


dram = {}
for name, c in summ.items():
    short = name.split(" #")[0]
    short = short if "chain_eval" in short else short.split("<")[0]
    v = num(c["dram__bytes_read.sum"]) + num(c["dram__bytes_write.sum"])
    dram[short] = v if v == v else None   # ncu reports nan for a launch too short to sample (the QR kernel on 0.2 % of the batch)
print(json.dumps({"workload": workload, "configs_per_step": B, "dram_bytes_per_launch": dram,
                  "command": "ncu --set full --clock-control none -s 4 -c 4 (the timed step's four kernels) python bench.py --steps 1 --warmup 1 "
                             "--no-extra --no-cpu-baseline", "counters": summ}, indent=1))

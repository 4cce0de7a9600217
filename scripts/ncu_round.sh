#!/bin/bash
# usage (on the GPU box): scripts/ncu_round.sh r02
# One `ncu --set full` capture of the dominant kernel of every bench workload, each after the same command line has
# exited 0 without ncu (B200_PROFILING.md).  Reports land in gpurun_out/ncu_<round>_<name>.ncu-rep; summarise them at
# home with scripts/ncu_summary.py into profiles/.  ONLY="name name" restricts the captures.
round=${1:-r02}
common="--no-e2e --no-cpu --no-verify --configs none --warmup 3"
run() {  # name  kernel-regex  skip  env  bench args...
  local name=$1 regex=$2 skip=$3 envs=$4; shift 4
  if [ -n "$ONLY" ]; then case " $ONLY " in *" $name "*) ;; *) return ;; esac; fi
  env $envs python bench.py "$@" $common > gpurun_out/ncu_${round}_${name}.plain.log 2>&1 &&
  env $envs ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
      -o gpurun_out/ncu_${round}_${name} -f python bench.py "$@" $common > gpurun_out/ncu_${round}_${name}.log 2>&1
  echo "$name rc=$?"
  # the reports are too large to bring home all at once: digest them here (KEEP_REPS="name name" keeps some)
  python scripts/ncu_summary.py gpurun_out/ncu_${round}_${name}.ncu-rep > gpurun_out/ncu_${round}_${name}.txt 2>&1
  case " $KEEP_REPS " in *" $name "*) ;; *) rm -f gpurun_out/ncu_${round}_${name}.ncu-rep ;; esac
  rm -f gpurun_out/ncu_${round}_${name}.plain.log gpurun_out/ncu_${round}_${name}.log
}
run vol3d_tb2   vol3d_tb2        2 X=1                 --workload heat3d --steps 4
run vol3d_star7 'vol3d_kernel'   4 MASSIV_B200_TB2=0   --workload heat3d --steps 4
run sobel       tile2d_kernel    8 X=1                 --workload sobel --steps 4
run gauss7      tile2d_kernel    4 X=1                 --workload gauss7 --steps 4
run gauss7sep   sep2d_kernel     4 X=1                 --workload gauss7sep --steps 4
run avg3x3      tile2d_kernel    4 X=1                 --workload avg3x3 --steps 4
run life        life_bits        1 X=1                 --workload life --steps 1000
run sobel_closure smallwin       4 X=1                 --workload sobel_closure --steps 4
run sum9x9_u8   smallwin         4 X=1                 --workload sum9x9_u8 --steps 4
run avg3d       smallwin         4 X=1                 --workload avg3d --steps 4
run tb2_odd     vol3d_tb2        2 X=1                 --workload heat3d --grid 1535 1535 1535 --steps 4
run sobel_odd   tile2d_kernel    4 X=1                 --workload sobel --grid 32767 32767 --steps 4

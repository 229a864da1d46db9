# Same-box A/B of two builds of the library: the tree's libmmw_radar_b200.so against profiles/ab/libmmw_base.so (a build of
# an earlier commit, made by hand and git-ignored), through profiles/stage_times.py.  usage: bash profiles/tools/ab_lib.sh TAG [workloads...]
set -x
cd $GRAFT_REPO_ROOT
TAG=$1; shift
PKG=cuda-based-mmwave-radar-object-detection-acceleration_b200
mkdir -p gpurun_out
timeout 300 python profiles/stage_times.py "$@" > gpurun_out/stage_times_${TAG}_new.log 2>&1; echo new rc=$?
cp $PKG/libmmw_radar_b200.so /tmp/new.so
cp profiles/ab/libmmw_base.so $PKG/libmmw_radar_b200.so
timeout 300 python profiles/stage_times.py "$@" > gpurun_out/stage_times_${TAG}_base.log 2>&1; echo base rc=$?
cp /tmp/new.so $PKG/libmmw_radar_b200.so
timeout 300 python profiles/stage_times.py "$@" > gpurun_out/stage_times_${TAG}_new2.log 2>&1; echo new2 rc=$?
grep -h "keep=0" gpurun_out/stage_times_${TAG}_new.log gpurun_out/stage_times_${TAG}_base.log gpurun_out/stage_times_${TAG}_new2.log

# one compute-sanitizer tool per gpurun call (B200_PROFILING.md): bash tools/r2_sanitize.sh racecheck|synccheck|memcheck
TOOL=${1:-racecheck}
mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/r02_sanitize_plain_$TOOL.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_small.py > gpurun_out/r02_sanitize_$TOOL.log 2>&1
echo "exit $?"; tail -12 gpurun_out/r02_sanitize_$TOOL.log

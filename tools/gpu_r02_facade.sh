cd "${GRAFT_REPO_ROOT:-/root/repo}"
cat /sys/kernel/mm/transparent_hugepage/enabled
python - <<'PY'
from msufsort_b200 import textgen
textgen.markov3(1<<28).tofile('/dev/shm/t.bin')
PY
msufsort_b200/lib/facade_bench /dev/shm/t.bin 3 2
B200SA_COPY_THREADS=8 msufsort_b200/lib/facade_bench /dev/shm/t.bin 3 2
rm -f /dev/shm/t.bin

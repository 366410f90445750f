# The fused command three times on one file (the first CUDA processes on a fresh box pay the driver start-up).
# C2 x 0.4 (20 M reads) in /dev/shm, --gpu-deflate, level 1.  Measurement tool.
cd "$(dirname "$0")/../.."
python - <<'PY'
import sys
sys.path.insert(0, '.')
from openge_b200 import bamhost, bamio, synth
bam = synth.make("C2", 0.4, seed=2)
raw = bamio.serialize_bam_stream(bam)
open("/dev/shm/in.bam", "wb").write(bamhost.bgzf_compress(raw, 1))
PY
B=openge_b200/host/_build/oge_dedup_fused
$B dedup /dev/shm/in.bam -o /dev/shm/o.bam -v --nopg -c 1 --gpu-deflate > /dev/null 2>&1
for mode in "" "" ""; do
  echo "== mode [$mode]"
  s=$(date +%s.%N)
  $B dedup /dev/shm/in.bam -o /dev/shm/o.bam -v --nopg -c 1 --gpu-deflate $mode > /dev/shm/log.txt 2>&1
  e=$(date +%s.%N)
  grep -E "Timing|gpu deflate" /dev/shm/log.txt
  echo "wall $(python -c "print('%.3f' % ($e - $s))") s"
done
rm -f /dev/shm/in.bam /dev/shm/o.bam /dev/shm/log.txt

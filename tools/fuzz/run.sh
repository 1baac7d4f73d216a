#!/bin/sh
# Host loaders under AddressSanitizer + UBSan on mutated inputs (images, scene JSON, OBJ, config JSON, checkpoints).  CPU only.
#   sh tools/fuzz/run.sh [work_dir] [seed] [count]
# Prints one "ok N bad M" line per batch; any "runtime error" / "AddressSanitizer" line is a finding (exit status 1).
set -e
HERE=$(cd "$(dirname "$0")" && pwd); REPO=$(cd "$HERE/../.." && pwd); SRC=$REPO/pathtracerwithcuda_b200/csrc
WORK=${1:-/tmp/ptb_fuzz}; SEED=${2:-1}; COUNT=${3:-3000}
CXX=/usr/bin/g++; [ -x $CXX ] || CXX=g++
FLAGS="-std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -I$SRC -I$REPO/include"
mkdir -p "$WORK/bin"
$CXX $FLAGS "$HERE/fuzz_images.cpp" "$SRC/scene_io.cpp" "$SRC/jpeg_decode.cpp" -o "$WORK/bin/images"
$CXX $FLAGS -pthread "$HERE/fuzz_scene.cpp" "$SRC/scene_io.cpp" "$SRC/jpeg_decode.cpp" -o "$WORK/bin/scene"
$CXX $FLAGS "$HERE/fuzz_config.cpp" "$SRC/scene_io.cpp" "$SRC/jpeg_decode.cpp" -o "$WORK/bin/config"
$CXX $FLAGS "$HERE/fuzz_checkpoint.cpp" "$SRC/image_out.cpp" -o "$WORK/bin/checkpoint"
$CXX $FLAGS -fopenmp -pthread "$HERE/fuzz_bvh.cpp" "$SRC/scene_io.cpp" "$SRC/jpeg_decode.cpp" "$SRC/bvh_host.cpp" -o "$WORK/bin/bvh"
python "$HERE/make_corpus.py" "$WORK/corpus" "$SEED" "$COUNT"
LOG="$WORK/log.txt"; : > "$LOG"
(cd "$WORK/corpus/images" && ls | xargs -n 1000 "$WORK/bin/images") >> "$LOG" 2>&1 || true
(ls "$WORK/corpus/scene_root/fz" | sed "s#^#$WORK/corpus/scene_root/fz/#" | xargs -n 500 "$WORK/bin/scene" "$WORK/corpus/scene_root") >> "$LOG" 2>&1 || true
(ls "$WORK/corpus/scene_root/fz" | grep "^o" | sed "s#^#$WORK/corpus/scene_root/fz/#" | xargs -n 300 "$WORK/bin/bvh" "$WORK/corpus/scene_root") >> "$LOG" 2>&1 || true   # host SAH builder, flatten, 8-wide collapse
(cd "$WORK/corpus/configs" && ls | xargs -n 1000 "$WORK/bin/config") >> "$LOG" 2>&1 || true
"$WORK/bin/checkpoint" >> "$LOG" 2>&1 || true
python "$HERE/oom_sweep.py" "$WORK/oom" >> "$LOG" 2>&1 || echo "exception oom sweep found abnormal exits" >> "$LOG"   # allocation / thread-start failures inside the threaded scene loader
grep "^ok" "$LOG"
if grep -q "runtime error\|AddressSanitizer\|exception" "$LOG"; then grep "runtime error\|SUMMARY\|exception" "$LOG" | sort | uniq -c | head -20; exit 1; fi
echo "no findings"

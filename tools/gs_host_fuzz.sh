#!/bin/bash
# Long fuzz of the HOST build of the GPU self-play kernels (tests/native/gs_host.cpp; no GPU needed): whole self-play runs
# with a pseudo-network over random seeds and configurations (simulations, mate depth, tree reuse, noise, game length
# limit, prior sharpness, playout-cap oscillation), each on three builds -- ASan+UBSan with small pools (suspension and
# record truncation all the time), ASan+UBSan with the real pool sizes, and an optimised build with more games.  Every
# run audits every tree after every round and replays every finished game; the script stops at the first failure.
#   tools/gs_host_fuzz.sh <seconds> [work dir] [first seed]
set -u
SECS=${1:-600}; DIR=${2:-/tmp/gsfuzz}; SEED0=${3:-1000}; ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$DIR"
INC="-Wno-unknown-pragmas -Wno-attributes -I$ROOT/tests/native/shim -I$ROOT/erweitern_55_b200/csrc"
SAN="g++ -O1 -g -std=c++20 -ffp-contract=off -fsanitize=address,undefined -fno-sanitize-recover=all -fno-omit-frame-pointer $INC"
$SAN -o "$DIR/san" "$ROOT/tests/native/gs_host.cpp" || exit 2
$SAN -DE55_GS_MAX_NODES=512 -DE55_GS_MAX_EDGES=12288 -DE55_GS_RECORD_WORDS=1024 -o "$DIR/san_small" "$ROOT/tests/native/gs_host.cpp" || exit 2
g++ -O2 -std=c++20 -ffp-contract=off $INC -o "$DIR/fast" "$ROOT/tests/native/gs_host.cpp" || exit 2
pick() { python3 -c "import random,sys;random.seed($1);print(random.choice(sys.argv[1:]))" "${@:2}"; }
T0=$(date +%s); i=0; runs=0
while [ $(( $(date +%s) - T0 )) -lt "$SECS" ]; do
  i=$((i+1)); seed=$((SEED0+i))
  sharp=$(pick $seed 0.0 1.0 3.0 6.0 12.0 30.0 100.0); sims=$(pick $((seed+1)) 16 32 160 480 800 1024)
  mate=$(pick $((seed+2)) 0 3 5 7); maxm=$(pick $((seed+3)) 30 80 512 512)
  reuse=$((i%3!=0)); noise=$((i%2)); cap=$((i%5==0))
  for bin in san_small san fast; do
    games=16; moves=800; [ $bin = fast ] && games=64 && moves=6000
    out=$("$DIR/$bin" $games $moves $sims $mate $reuse $noise $maxm $seed $sharp $cap 800 128 0.25 64 - 2>&1); rc=$?
    runs=$((runs+1))
    echo "$i $bin sims=$sims mate=$mate reuse=$reuse noise=$noise max=$maxm sharp=$sharp cap=$cap rc=$rc $(echo "$out" | tail -1)" >> "$DIR/log.txt"
    if [ $rc -ne 0 ]; then echo "FAILED: $bin seed $seed"; echo "$out" | tail -5; exit 1; fi
  done
done
echo "done: $runs runs ($i configurations x 3 builds), no failure"

#!/bin/bash
# Sweep of the Sudoku pipeline knobs at shard sizes (multi-GPU shards are 125k-500k puzzles).
for n in 125000 500000; do
  for fb in 1024 2048 4096 8192; do
    echo "== n=$n FIRST_BUDGET=$fb"
    DQ_TRACE=1 DQ_SUDOKU_FIRST_BUDGET=$fb python scripts/sudoku_bench.py $n 2>&1 | tail -2
  done
  for dm in 16 24 96; do
    echo "== n=$n DONATE_MIN=$dm"
    DQ_TRACE=1 DQ_SUDOKU_DONATE_MIN=$dm python scripts/sudoku_bench.py $n 2>&1 | tail -2
  done
  for dg in 32 48 192; do
    echo "== n=$n DONATE_GAP=$dg"
    DQ_TRACE=1 DQ_SUDOKU_DONATE_GAP=$dg python scripts/sudoku_bench.py $n 2>&1 | tail -2
  done
done

# how much do wave quantisation + per-launch head/tail cost on the res4 layers?  batch 8 (256 M tiles, 1.73 rounds) vs
# batch 37 (1184 = 8 x 148 tiles) vs batch 74
for b in 8 37 74; do echo "batch $b"; timeout 300 python tools/bench_conv_layers.py --batch $b --only "res4 2,res3 2b,res5 2b" --reps 7 2>&1 | grep -v "^sum\|^layer"; done

# sustained (power-capped) timing of self-attention kernel variants: 4000 back-to-back launches each
export AGENDA_KNOBS=1
for v in 60 62 63 64 20 23 13 14; do
  echo -n "variant $v: "; REPS=4000 VARIANT=$v timeout 60 tools/ubench/bench_self.bin 0 | head -1
done
echo -n "unit v2: "; REPS=4000 timeout 60 tools/ubench/bench_self.bin 1 | head -1
echo -n "unit v3 ks2: "; V3=1 REPS=4000 timeout 60 tools/ubench/bench_self.bin 1 | head -1
echo -n "unit v3 ks2 emu4: "; V3=1 AGENDA_V3_EMU=4 REPS=4000 timeout 60 tools/ubench/bench_self.bin 1 | head -1
echo -n "unit v3 ks1: "; V3=1 AGENDA_V3_EMU=13 REPS=4000 timeout 60 tools/ubench/bench_self.bin 1 | head -1

for c in c2 c3 c4; do
for g in 32,1 64,1 128,1 176,1 256,1 64,2 128,2 176,2 256,2 128,3 256,3 256,4 128,4; do
  echo -n "$c $g : "; MCD_GEOMETRY=$g python tools/profile_config.py $c --calls 200 2>/dev/null | sed 's/.*| \([0-9.]* us per call\) | \(grid [0-9x ]*\),.*/\1 \2/'
done; done

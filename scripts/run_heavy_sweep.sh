for f in fre-nctools_b200/variants/*.so; do
  echo "== $f"
  XGRID_B200_LIB=$PWD/$f python scripts/profile_rank.py 0 8 2>/dev/null
  XGRID_B200_LIB=$PWD/$f python scripts/profile_rank.py 3 8 2>/dev/null
  XGRID_B200_LIB=$PWD/$f python scripts/clip_variants.py child 2 2>/dev/null
done

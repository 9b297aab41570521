cd $GRAFT_REPO_ROOT
for sc in hall rand cubes; do
  python profiles/stats.py $sc 2>&1 | tail -1
  SRB_NO_BLOCK_REJECT=1 python profiles/stats.py $sc 2>&1 | tail -1
done

# usage: bash profiles/run_ab2.sh "<scene list>" <frames> "<knobs A>" ...
cd $GRAFT_REPO_ROOT
scenes=$1; frames=$2; shift 2
for sc in $scenes; do for cfg in "$@"; do
  if [ "$cfg" = "-" ]; then env timeout 300 python profiles/ab.py $sc $frames 2>&1 | tail -1 | cut -c1-420
  else env $cfg timeout 300 python profiles/ab.py $sc $frames 2>&1 | tail -1 | cut -c1-420; fi
done; done

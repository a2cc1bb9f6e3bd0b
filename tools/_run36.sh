cd $GRAFT_REPO_ROOT
echo device; python tools/_run36.py 2>&1 | tail -10 | grep -o "overhead.*" | tr "\n" " "; echo
echo host; C3POA_HOST_ORDER=1 python tools/_run36.py 2>&1 | tail -10 | grep -o "overhead.*" | tr "\n" " "; echo

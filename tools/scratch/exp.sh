for d in 8 2; do echo "=== DFB head dbg bits $d"; DFB_HEAD_DBG=$d timeout 300 python tools/timeline.py 2>&1 | sed -n '/== head/,$p' | grep -E "mma2 issued|D2 ready|epi end|L3 issued"; done

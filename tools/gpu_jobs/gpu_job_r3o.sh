SAN_CASES="c2pad c2pad60 c2big c2hast c4rw c4pad" tools/run_sanitizers.sh r2f

python scripts/r02_dbg_multi.py 2>&1 | tail -20
echo "--- system nccl"
LMB200_NCCL_LIB=/usr/lib/x86_64-linux-gnu/libnccl.so.2 python scripts/r02_dbg_multi.py 2>&1 | tail -20

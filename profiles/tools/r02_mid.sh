for n in 8192; do
  for k in auto duo quartet; do
    echo -n "$n $k: "; OALSFX_KERNEL=$k python profiles/tools/r02_one.py chain $n 2>&1 | grep -o "'ms_per_block_device': [0-9.]*"
  done
  echo -n "$n span-bulk-forced: "; OALSFX_SPAN_BULK=2 OALSFX_KERNEL=span python profiles/tools/r02_one.py chain $n 2>&1 | grep -o "'ms_per_block_device': [0-9.]*"
done

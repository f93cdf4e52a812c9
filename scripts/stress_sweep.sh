for o in cta_val_limit=200 cta_val_limit=400 cta_val_limit=700 cta_node_limit=100 cta_node_limit=300 cta_node_limit=700 cta_val_limit=300,cta_no_reserve=1 cta_node_limit=300,cta_no_reserve=1 cta_val_limit=500,cta_no_reserve=2 cta_node_limit=400,cta_no_reserve=2 cta_item_limit=40 cta_item_limit=100,cta_val_limit=600 sums_ask=1 sums_ask=1,cta_val_limit=300; do
  echo "== $o"; SWEEP_OPTS=$o timeout 200 python scripts/parity_sweep.py 42 5 9 300 2>&1 | tail -n 2
done

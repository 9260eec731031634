#!/bin/bash
# time the SVC engine kernel of experiment builds (variants/*.so swapped in for libtcsdn.so); results are NOT valid labels
cp traffic_classifier_sdn_b200/libtcsdn.so /tmp/libtcsdn_product.so
for v in product "$@"; do
  if [ "$v" != product ]; then cp variants/libtcsdn_$v.so traffic_classifier_sdn_b200/libtcsdn.so; fi
  echo "== $v"
  bash tools/gpu_svc_time.sh 2>&1 | grep engine_kernel | head -1
done
cp /tmp/libtcsdn_product.so traffic_classifier_sdn_b200/libtcsdn.so

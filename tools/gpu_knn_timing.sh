#!/bin/bash
# cycle breakdown of the KNN engine's three roles (experiment build variants/libtcsdn_knntiming.so, -DTCSDN_EXP_KNN_TIMING)
cp traffic_classifier_sdn_b200/libtcsdn.so /tmp/libtcsdn_product.so
cp variants/libtcsdn_knntiming.so traffic_classifier_sdn_b200/libtcsdn.so
timeout 300 python tools/run_workload.py knn 10000000 1 2>&1 | tail -5
cp /tmp/libtcsdn_product.so traffic_classifier_sdn_b200/libtcsdn.so

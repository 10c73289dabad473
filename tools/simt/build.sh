#!/bin/bash
# Host build of the device sources against the SIMT interpreter (tools/simt/simt.h); the tests build it on demand too.
set -e
cd "$(dirname "$0")"
g++ -std=c++17 -O1 -g -fPIC -shared -Wno-unknown-pragmas sim_compact.cpp -o libpgs_sim.so

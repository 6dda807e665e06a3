#!/bin/bash
# Registers / stack (spills) / shared memory of every kernel in the built library (cuobjdump resource usage).
lib=${1:-bvh_b200/libbvh_c.so}
cuobjdump --dump-resource-usage "$lib" 2>/dev/null | awk '/^ *Function /{name=$2; sub(/:$/,"",name); getline; print name, $0}' | while read name rest; do
  short=$(echo "$name" | c++filt | sed -E 's/bvhb200::\(anonymous namespace\):://g; s/\(bvhb200.*//; s/^void //')
  echo "$short | $(echo $rest | sed -E 's/CONSTANT\[[0-9]\]:[0-9]+ ?//g; s/TEXTURE:0 ?//; s/SURFACE:0 ?//; s/SAMPLER:0 ?//')"
done

#!/bin/bash
# parameter sweep of the persistent-trace constants (run on the GPU box; rebuilds ddgi.o each time)
cd vulkanexp_b200/csrc
for refill in 4 8 16; do for chunk in 32 64 128; do
  touch ddgi.cu; make -s -j8 EXTRA="-DPT_REFILL_MIN=$refill -DPT_CHUNK=${chunk}u" > /dev/null 2>&1
  echo "refill=$refill chunk=$chunk $(cd ../..; python tools/profile_step.py 4 | tail -1 | sed 's/.*trace_primary/trace_primary/')"
done; done

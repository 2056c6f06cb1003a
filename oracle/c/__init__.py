"""C restatement of the reference's serial u64 CPU path (see ref_u64.c).  TEST INFRASTRUCTURE / CPU BASELINE."""

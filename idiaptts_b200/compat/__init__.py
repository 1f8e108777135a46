"""pyworld- / pysptk-compatible call signatures backed by libb200world.so (numpy in, numpy out)."""

/* Stand-in for Maxeler's MaxSLiCInterface.h (TRPOCpuCode.c:12); see Maxfiles.h in this directory. Empty on purpose. */

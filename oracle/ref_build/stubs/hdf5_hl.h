/* empty: stand-in for <hdf5_hl.h>; see hdf5.h here */

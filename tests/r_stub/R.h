/* stand-in for R.h (see Rinternals.h in this directory) */
#include <stdlib.h>

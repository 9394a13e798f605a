/* Scene.h:8 includes "MTRand.h"; the file in the tree is mtrand.h (case-insensitive FS upstream). */
#include "mtrand.h"

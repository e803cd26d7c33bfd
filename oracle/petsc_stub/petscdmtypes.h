#include "petsc_stub.h"

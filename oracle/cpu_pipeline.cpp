#include "vm_oracle.h"

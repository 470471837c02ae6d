#pragma once
#include "lammps.h"

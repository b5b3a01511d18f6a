// TEST INFRASTRUCTURE — see task_group.h (scheduling stand-in, no arithmetic).
#pragma once
#include "task_group.h"

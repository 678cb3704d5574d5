#pragma once
#define LTETRIGGER_API

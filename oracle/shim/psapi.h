/* empty stand-in for <psapi.h> (nothing on the hot path uses it) */
#pragma once

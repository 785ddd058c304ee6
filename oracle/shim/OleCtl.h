// Build scaffolding for oracle/_ref ONLY: Win32 GUID stubs (Utilities.cpp:23,288-290).
#pragma once
struct GUID { unsigned long Data1; unsigned short Data2; unsigned short Data3; unsigned char Data4[8]; };
#define S_OK 0
inline int CoCreateGuid(GUID* g) { *g = GUID{}; return S_OK; }

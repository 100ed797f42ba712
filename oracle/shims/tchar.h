// stand-in for MSVC <tchar.h> (STDAFX:14); the reference's entry point becomes ref_tmain
#pragma once
typedef char _TCHAR;
#define _tmain ref_tmain

#pragma once
// Export macro of the public API (the reference's ML/dll.hpp:8-20: empty outside Windows).
#if defined(_WIN32)
#ifdef ML_EXPORTS
#define DLL_DECLSPEC __declspec(dllexport)
#else
#define DLL_DECLSPEC __declspec(dllimport)
#endif
#else
#define DLL_DECLSPEC
#endif

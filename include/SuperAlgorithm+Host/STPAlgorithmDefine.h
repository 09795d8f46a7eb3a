// Export macro of the SuperAlgorithm+Host library (SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPAlgorithmDefine.h:5-13).
#pragma once
#if defined(_WIN32)
#ifdef SUPERALGORITHMPLUS_HOST_EXPORTS
#define STP_ALGORITHM_HOST_API __declspec(dllexport)
#else
#define STP_ALGORITHM_HOST_API __declspec(dllimport)
#endif
#else
#define STP_ALGORITHM_HOST_API __attribute__((visibility("default")))
#endif

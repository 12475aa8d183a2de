/* Test-infrastructure shim (NOT product code, NOT an OpenCL implementation): the minimum
 * typedefs/constants/prototypes that Code/openCL_tool.h needs to *compile* so that
 * Code/Convex_Tree.h:16 can be included by the CPU oracle harness. No OpenCL call is ever made. */
#ifndef SCHT_SHIM_CL_H
#define SCHT_SHIM_CL_H
#include <stddef.h>
#include <stdint.h>
typedef int32_t cl_int; typedef uint32_t cl_uint; typedef int64_t cl_long; typedef uint64_t cl_ulong;
typedef cl_uint cl_bool; typedef cl_ulong cl_bitfield; typedef cl_bitfield cl_device_type;
typedef cl_uint cl_platform_info; typedef cl_uint cl_device_info; typedef cl_uint cl_program_build_info;
typedef cl_bitfield cl_mem_flags; typedef cl_bitfield cl_command_queue_properties;
typedef struct _cl_platform_id* cl_platform_id; typedef struct _cl_device_id* cl_device_id;
typedef struct _cl_context* cl_context; typedef struct _cl_command_queue* cl_command_queue;
typedef struct _cl_mem* cl_mem; typedef struct _cl_program* cl_program;
typedef struct _cl_kernel* cl_kernel; typedef struct _cl_event* cl_event;
typedef intptr_t cl_context_properties;
#define CL_SUCCESS 0
#define CL_TRUE 1
#define CL_FALSE 0
#define CL_DEVICE_TYPE_DEFAULT (1 << 0)
#define CL_DEVICE_TYPE_CPU (1 << 1)
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_DEVICE_TYPE_ACCELERATOR (1 << 3)
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PROGRAM_BUILD_STATUS 0x1181
#define CL_PLATFORM_NAME 0x0902
#define CL_PLATFORM_VENDOR 0x0903
#define CL_PLATFORM_VERSION 0x0901
#define CL_PLATFORM_PROFILE 0x0900
#define CL_PLATFORM_EXTENSIONS 0x0904
#define CL_DEVICE_NAME 0x102B
#define CL_DEVICE_VENDOR 0x102C
#define CL_DEVICE_VERSION 0x102F
#define CL_DEVICE_MAX_COMPUTE_UNITS 0x1002
#define CL_DEVICE_MAX_WORK_GROUP_SIZE 0x1004
#define CL_DEVICE_GLOBAL_MEM_SIZE 0x101F
#define CL_DEVICE_LOCAL_MEM_SIZE 0x1023
#define CL_DEVICE_EXTENSIONS 0x1030
#define CL_CONTEXT_PLATFORM 0x1084
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#ifdef __cplusplus
extern "C" {
#endif
cl_int clGetPlatformIDs(cl_uint, cl_platform_id*, cl_uint*);
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void*, size_t*);
cl_int clGetDeviceIDs(cl_platform_id, cl_device_type, cl_uint, cl_device_id*, cl_uint*);
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void*, size_t*);
cl_context clCreateContext(const cl_context_properties*, cl_uint, const cl_device_id*, void (*)(const char*, const void*, size_t, void*), void*, cl_int*);
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int*);
cl_int clReleaseMemObject(cl_mem);
cl_int clReleaseProgram(cl_program);
cl_program clCreateProgramWithSource(cl_context, cl_uint, const char**, const size_t*, cl_int*);
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id*, const char*, void (*)(cl_program, void*), void*);
cl_int clGetProgramBuildInfo(cl_program, cl_device_id, cl_program_build_info, size_t, void*, size_t*);
cl_kernel clCreateKernel(cl_program, const char*, cl_int*);
#ifdef __cplusplus
}
#endif
#endif

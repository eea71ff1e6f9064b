/* Minimal OpenCL 1.2 declarations (TEST INFRASTRUCTURE): just enough for the reference's src/main.cpp to COMPILE its
 * CUDA flavour -- main.cpp includes opencl_manager.h unconditionally, while this image ships no CL/cl.h.  Nothing
 * here is ever called or linked: with -DCUDA_DECODER main.cpp instantiates only ldpc_decoder_gpu_cuda. */
#ifndef LDPC_B200_CL_STUB_H
#define LDPC_B200_CL_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef uint8_t cl_uchar;
typedef cl_uint cl_bool;
typedef union { cl_uchar s[4]; struct { cl_uchar x, y, z, w; }; } cl_uchar4;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_map_flags;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_command_queue_properties;
typedef cl_uint cl_device_info;
typedef cl_uint cl_platform_info;
typedef cl_uint cl_program_build_info;
typedef cl_uint cl_kernel_work_group_info;
typedef intptr_t cl_context_properties;
typedef struct _cl_platform_id *cl_platform_id;
typedef struct _cl_device_id *cl_device_id;
typedef struct _cl_context *cl_context;
typedef struct _cl_command_queue *cl_command_queue;
typedef struct _cl_mem *cl_mem;
typedef struct _cl_program *cl_program;
typedef struct _cl_kernel *cl_kernel;
typedef struct _cl_event *cl_event;
#define CL_SUCCESS 0
#define CL_TRUE 1
#define CL_FALSE 0
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_USE_HOST_PTR (1 << 3)
#define CL_MEM_ALLOC_HOST_PTR (1 << 4)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_MAP_READ (1 << 0)
#define CL_MAP_WRITE (1 << 1)
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *);
void *clEnqueueMapBuffer(cl_command_queue, cl_mem, cl_bool, cl_map_flags, size_t, size_t, cl_uint, const cl_event *,
                         cl_event *, cl_int *);
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, const void *, cl_uint, const cl_event *,
                            cl_event *);
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void *, cl_uint, const cl_event *,
                           cl_event *);
cl_int clEnqueueFillBuffer(cl_command_queue, cl_mem, const void *, size_t, size_t, size_t, cl_uint, const cl_event *,
                           cl_event *);
#ifdef __cplusplus
}
#endif
#endif

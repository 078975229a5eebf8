/* ggml-b200.h — backend-level C ABI of libggml_b200.so.
 *
 * It exports EXACTLY the symbols of the reference's ggml/include/ggml-cuda.h:24-46 (plus ggml_backend_cuda_reg_devices,
 * ggml/src/ggml-cuda.cu:5520), so a reference build configured with GGML_USE_CUDA links against it instead of its own
 * ggml-cuda.cu and llama-bench / llama-server drive it unchanged through ggml_backend_graph_compute.
 * The opaque handle types are ggml's own (ggml/include/ggml-backend.h); this header only repeats the prototypes so that the
 * boundary is documented in this repository.  Compiled against the reference headers at build time (never copied).
 */
#ifndef GGML_B200_H
#define GGML_B200_H
#include <stddef.h>
#include <stdbool.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ggml_backend * ggml_backend_t;
typedef struct ggml_backend_buffer_type * ggml_backend_buffer_type_t;
typedef void (*ggml_log_callback_b200)(int level, const char * text, void * user_data);

ggml_backend_t             ggml_backend_cuda_init(int device, const void * params /* "k=v,..." or NULL */, const void * model); /* ggml-cuda.h:24 */
bool                       ggml_backend_is_cuda(ggml_backend_t backend);                                                          /* :26 */
ggml_backend_buffer_type_t ggml_backend_cuda_buffer_type(int device);                                                             /* :29 */
ggml_backend_buffer_type_t ggml_backend_cuda_split_buffer_type(const float * tensor_split);                                       /* :32 */
ggml_backend_buffer_type_t ggml_backend_cuda_host_buffer_type(void);                                                              /* :35 */
int                        ggml_backend_cuda_get_device_count(void);                                                              /* :37 */
void                       ggml_backend_cuda_get_device_description(int device, char * description, size_t description_size);    /* :38 */
void                       ggml_backend_cuda_get_device_memory(int device, size_t * free, size_t * total);                       /* :39 */
bool                       ggml_backend_cuda_register_host_buffer(void * buffer, size_t size);                                   /* :41 */
void                       ggml_backend_cuda_unregister_host_buffer(void * buffer);                                              /* :42 */
void                       ggml_backend_cuda_log_set_callback(ggml_log_callback_b200 log_callback, void * user_data);            /* :44 */
void                       ggml_backend_cuda_invalidate_graphs(const void * model);                                              /* :46 */
int                        ggml_backend_cuda_reg_devices(void);                                                                   /* ggml-cuda.cu:5520 */
#ifdef __cplusplus
}
#endif
#endif

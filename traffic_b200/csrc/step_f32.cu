// fp32 ("throughput mode") instantiation of the fused step kernel: state in coordinates relative to the map origin,
// predicates evaluated in fp32 only (see step_kernel.cuh).  FMA contraction is allowed here.
#include "step_kernel.cuh"
#include "resident_kernel.cuh"

void launch_step(const StepArgs32& a, cudaStream_t stream) { tb2k::launch_step_t<float>(a, stream); }

int persistent_grid_capacity_f32() { return tb2k::persistent_grid_capacity_t<float>(); }

cudaError_t launch_persistent(const StepArgs32& a, const PersistArgsT<float>& pa, cudaStream_t stream) {
    return tb2k::launch_persistent_t<float>(a, pa, stream);
}

void launch_prepare(const StepArgs32& a, float2* pt_curv, float2* face_unit, int n_faces, cudaStream_t stream) {
    tb2k::launch_prepare_t<float>(a, pt_curv, face_unit, n_faces, stream);
}

cudaError_t launch_resident(const StepArgs32& a, const PersistArgsT<float>& pa, cudaStream_t stream) {
    return tb2k::launch_resident_t<float>(a, pa, stream);
}

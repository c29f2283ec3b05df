#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/).

Builds the UNMODIFIED-ALGORITHM reference (DavidXu-JJ/PoissonRecon_GPU, /root/reference) for
sm_100 so it can be run on the B200 box next to this repo's pipeline:

  oracle/_ref/ref_tables          reference host-side B-spline code (g++, CPU)     -> golden tables
  oracle/_ref/ref_poisson_d<D>    reference main.cu pipeline, maxDepth=D (nvcc)    -> parity + baseline

Nothing from the reference is copied into the repository: sources are read where they lie,
a *harness patch* is applied to a scratch copy under /tmp, and only binaries land in the
git-ignored oracle/_ref/.  The patch does not touch any kernel or algorithm; it only
  (1) takes input / output paths from argv instead of main.cu:3251-3252,
  (2) takes maxDepth from -DREF_DEPTH instead of main.cu:69,
  (3) dumps intermediate arrays to $REF_DUMP_DIR when that variable is set,
  (4) records the CG iteration count per depth (one extra store by thread 0 at kernel exit),
  (6) adds __launch_bounds__(1024) to the six kernels the reference launches with 1024-thread
      blocks (they exceed 64 registers on sm_100 with nvcc 12.9 and would silently fail to launch),
  (5) initialises two host ints the reference leaves uninitialised when a refinement pass has
      zero roots (main.cu:4439-4446, 4521-4527: cudaMemcpy from a null device pointer fails
      silently because CHECK is compiled out, Debug.cuh:8) -- without this the reference reads
      stack garbage as a vertex count.
The reference is only valid for 5 <= D <= 9.  --depths 10 additionally applies
patch_widen() ("ref+widen", BASELINE.md 2.1: the packed function index becomes a long long) and
writes oracle/_ref/ref_poisson_d10_widen; results of that binary are always labelled "ref+widen".

Usage: python oracle/build_ref.py [--depths 8 9] [--tables-only]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

DUMP_HELPERS = r'''
// ---- harness patch: dump helpers (not part of the reference) ----
static const char* refDumpDir(){ return getenv("REF_DUMP_DIR"); }
static void refDumpHost(const char* name,const void* h,size_t bytes){
    const char* d=refDumpDir(); if(!d) return;
    char path[1024]; snprintf(path,sizeof(path),"%s/%s.bin",d,name);
    FILE* fp=fopen(path,"wb"); if(!fp) return; fwrite(h,1,bytes,fp); fclose(fp);
}
static void refDumpDev(const char* name,const void* dptr,size_t bytes){
    const char* d=refDumpDir(); if(!d) return;
    void* h=malloc(bytes?bytes:1); cudaMemcpy(h,dptr,bytes,cudaMemcpyDefault);
    refDumpHost(name,h,bytes); free(h);
}
static void refLogPass(const char* kind,int nv,int nt){
    const char* d=refDumpDir(); if(!d) return;
    char path[1024]; snprintf(path,sizeof(path),"%s/passes.txt",d);
    FILE* fp=fopen(path,"a"); if(!fp) return; fprintf(fp,"%s %d %d\n",kind,nv,nt); fclose(fp);
}
// ---- end harness patch ----
'''


def sub_once(s, old, new, what):
    if s.count(old) != 1:
        raise SystemExit(f"harness patch anchor not unique/found ({what}): count={s.count(old)}")
    return s.replace(old, new)


def patch_main(src):
    s = src
    s = sub_once(s, "#define maxDepth 8\n", "#define maxDepth REF_DEPTH\n", "depth macro")
    s = sub_once(s, "int main() {", DUMP_HELPERS + "\nint main(int argc,char **argv) {\n    if(argc<3){fprintf(stderr,\"usage: %s in out\\n\",argv[0]);return 2;}", "main signature")
    s = sub_once(s, '    char fileName[]="/home/davidxu/bunny.points.ply";\n', "    char *fileName=argv[1];\n", "input path")
    s = sub_once(s, '    char outName[]="/home/davidxu/bunny.ply";\n', "    char *outName=argv[2];\n", "output path")
    # (3) dumps
    s = sub_once(s, '    printf("NodeArray_sz:%d\\n",NodeArray_sz);\n',
                 '    printf("NodeArray_sz:%d\\n",NodeArray_sz);\n'
                 '    refDumpDev("nodearray",NodeArray,sizeof(OctNode)*(size_t)NodeArray_sz);\n'
                 '    refDumpHost("base",BaseAddressArray,sizeof(int)*(maxDepth_h+1));\n'
                 '    refDumpHost("count",NodeArrayCount_h,sizeof(int)*(maxDepth_h+1));\n'
                 '    refDumpDev("points",samplePoints_d,sizeof(float)*3*(size_t)count);\n'
                 '    refDumpDev("normals",sampleNormals_d,sizeof(float)*3*(size_t)count);\n'
                 '    refDumpDev("p2n",PointToNodeArrayD,sizeof(int)*(size_t)count);\n'
                 '    { float cs[4]={center.coords[0],center.coords[1],center.coords[2],scale}; refDumpHost("center_scale",cs,sizeof(cs)); }\n',
                 "nodearray dump")
    s = sub_once(s, "    cudaFree(VectorField);\n",
                 '    refDumpDev("vectorfield",VectorField,sizeof(float)*3*(size_t)NodeDNum);\n'
                 '    cudaDeviceSynchronize(); refDumpDev("divergence",Divergence,sizeof(float)*(size_t)NodeArray_sz);\n'
                 "    cudaFree(VectorField);\n", "divergence dump")
    s = sub_once(s, "    cudaFree(Divergence);\n",
                 '    cudaDeviceSynchronize(); refDumpDev("x",d_x,sizeof(float)*(size_t)NodeArray_sz);\n'
                 "    cudaFree(Divergence);\n", "x dump")
    s = sub_once(s, "    isoValue/=count;\n",
                 "    isoValue/=count;\n"
                 '    refDumpHost("iso",&isoValue,sizeof(float));\n'
                 '    refDumpDev("pointvalue",pointValue,sizeof(float)*(size_t)count);\n', "iso dump")
    s = sub_once(s, '    printf("Compute vertex implicit function value takes:%lfs\\n",mid9-mid_insert);\n',
                 '    printf("Compute vertex implicit function value takes:%lfs\\n",mid9-mid_insert);\n'
                 '    refDumpDev("vvalue",vvalue,sizeof(float)*(size_t)VertexArray_sz);\n'
                 '    refDumpDev("vertexarray",VertexArray,sizeof(VertexNode)*(size_t)VertexArray_sz);\n'
                 '    refDumpDev("edgearray",EdgeArray,sizeof(EdgeNode)*(size_t)EdgeArray_sz);\n'
                 '    { int sz[3]={VertexArray_sz,EdgeArray_sz,FaceArray_sz}; refDumpHost("vef_sizes",sz,sizeof(sz)); }\n',
                 "vvalue dump")
    s = sub_once(s, '    printf("SubdivideNum:%d\\n",SubdivideNum);\n',
                 '    printf("SubdivideNum:%d\\n",SubdivideNum);\n'
                 '    refDumpDev("subdividenode",SubdivideNode,sizeof(OctNode)*(size_t)SubdivideNum);\n'
                 '    refDumpDev("nodearray_after_mc",NodeArray,sizeof(OctNode)*(size_t)NodeArray_sz);\n',
                 "subdivide dump")
    # pass logging: every insertTriangle call site
    s = sub_once(s, "    insertTriangle(VertexBuffer,allVexNums,\n                   TriangleBuffer,allTriNums,\n                   mesh);\n",
                 '    refLogPass("main",allVexNums,allTriNums);\n'
                 "    insertTriangle(VertexBuffer,allVexNums,\n                   TriangleBuffer,allTriNums,\n                   mesh);\n", "main pass log")
    s = sub_once(s, "        insertTriangle(SubdivideVertexBuffer,SubdivideAllVexNums,\n",
                 '        refLogPass("coarse",SubdivideAllVexNums,SubdivideAllTriNums);\n'
                 "        insertTriangle(SubdivideVertexBuffer,SubdivideAllVexNums,\n", "coarse pass log")
    s = sub_once(s, "        insertTriangle(RebuildVertexBuffer, RebuildAllVexNums,\n",
                 '        refLogPass("finer",RebuildAllVexNums,RebuildAllTriNums);\n'
                 "        insertTriangle(RebuildVertexBuffer, RebuildAllVexNums,\n", "finer pass log")
    # (5) uninitialised host ints when a finer pass has zero roots
    s = sub_once(s, "        int RebuildLastVexAddr;\n        int RebuildLastVexNums;\n",
                 "        int RebuildLastVexAddr=0;\n        int RebuildLastVexNums=0;\n", "uninit vex")
    s = sub_once(s, "        int RebuildLastTriAddr;\n        int RebuildLastTriNums;\n",
                 "        int RebuildLastTriAddr=0;\n        int RebuildLastTriNums=0;\n", "uninit tri")
    # (6) kernels the reference launches with 1024-thread blocks (main.cu:3370-3446, 1231-1256):
    # nvcc 12.9 / sm_100 gives some of them > 64 registers, so the launch fails with
    # cudaErrorLaunchOutOfResources -- silently, because CHECK is compiled out (Debug.cuh:8) --
    # and the stale error aborts the next thrust call.  A launch bound is a register-allocation
    # hint only; no kernel code changes.
    for kname in ("computeVectorField(", "precomputeEncodedFunctionIdxOfNode(", "computeEncodedFinerNodesDivergence(",
                  "generateDIdxArray(", "computeEncodedCoarserNodesDivergence(", "GenerateSingleNodeLaplacian("):
        tag = "__global__ void " + kname
        if s.count(tag) < 1:
            raise SystemExit("launch-bound anchor missing: " + kname)
        s = s.replace(tag, "__global__ void __launch_bounds__(1024) " + kname)
    s = sub_once(s, "int main(int argc,char **argv) {\n", "int main(int argc,char **argv) {\n    setvbuf(stdout,NULL,_IOLBF,0);\n", "line-buffered stdout")
    # final mesh dump
    s = sub_once(s, "    PlyWriteTriangles(outName,&mesh, PLY_ASCII,center,scale,NULL,0);\n",
                 '    if(mesh.inCorePoints.size()) refDumpHost("mesh_v",mesh.inCorePoints.data(),sizeof(float)*3*mesh.inCorePoints.size());\n'
                 '    { int nt=mesh.triangleCount(); std::vector<int> tt(3*(size_t)nt); mesh.resetIterator(); TriangleIndex ti; int fl;\n'
                 '      for(int q=0;q<nt;++q){ mesh.nextTriangle(ti,fl); tt[3*q]=ti.idx[0]; tt[3*q+1]=ti.idx[1]; tt[3*q+2]=ti.idx[2]; }\n'
                 '      refDumpHost("mesh_t",tt.data(),sizeof(int)*tt.size()); mesh.resetIterator(); }\n'
                 "    PlyWriteTriangles(outName,&mesh, PLY_ASCII,center,scale,NULL,0);\n", "mesh dump")
    return s


def patch_widen(src):
    """'ref+widen' (depth 10 only): the packed per-axis function index
    fi_x + fi_y*2^(D+1) + fi_z*2^(2(D+1)) (main.cu:889-899) needs 33 bits at maxDepth 10, so the
    reference's int overflows.  This patch only changes the TYPE of that packed index (and of the
    two decode divisors) to long long at its encode site and at every decode site
    (main.cu:1018-1040, 1107-1124, 1173-1197, 1344-1360, 2268-2311, 2339-2364, 2397-2422); no
    arithmetic, kernel structure or launch configuration changes."""
    s = src
    n_total = 0

    def rep(pattern, repl, what, minimum=1):
        nonlocal s, n_total
        s, n = re.subn(pattern, repl, s)
        if n < minimum:
            raise SystemExit(f"widen patch anchor missing ({what}): {n} < {minimum}")
        n_total += n
    rep(r"int \*EncodedNodeIdxInFunction", "long long *EncodedNodeIdxInFunction", "EncodedNodeIdxInFunction decls", 8)
    # (the unencoded precomputeFunctionIdxOfNode, main.cu:975, keeps its int[3] output)
    rep(r"__global__ void (__launch_bounds__\(1024\) )?precomputeEncodedFunctionIdxOfNode\(int \*BaseAddressArray_d,OctNode \*NodeArray,int NodeArray_sz,int \*NodeIdxInFunction\)",
        r"__global__ void \1precomputeEncodedFunctionIdxOfNode(int *BaseAddressArray_d,OctNode *NodeArray,int NodeArray_sz,long long *NodeIdxInFunction)", "encode kernel 1")
    rep(r"int \*DepthBuffer,int \*NodeIdxInFunction\)", "int *DepthBuffer,long long *NodeIdxInFunction)", "encode kernel 2")
    rep(r"(\n\s*)int \*NodeIdxInFunction,(\n)", r"\1long long *NodeIdxInFunction,\2", "coarse divergence param")
    rep(r"\(int \*\*\)&EncodedNodeIdxInFunction", "(long long **)&EncodedNodeIdxInFunction", "malloc cast")
    rep(r"int encode_idx", "long long encode_idx", "encode_idx locals", 6)
    rep(r"int decode_offset1=\(1<<\(maxD\+1\)\);", "long long decode_offset1=(1ll<<(maxD+1));", "decode_offset1", 6)
    rep(r"int decode_offset2=\(1<<\(2\*\(maxD\+1\)\)\);", "long long decode_offset2=(1ll<<(2*(maxD+1)));", "decode_offset2", 6)
    rep(r"void getEncodedFunctionIdxOfNode\(const int& key,const int &depthD,int \*idx\)",
        "void getEncodedFunctionIdxOfNode(const int& key,const int &depthD,long long *idx)", "encode signature")
    rep(r"\*idx = \(\(1<<depthD\)-1\)\*\(1\+\(1<<\(maxDepth\+1\)\)\+\(1<<\(2\*\(maxDepth\+1\)\)\) \);",
        "*idx = ((1ll<<depthD)-1)*(1+(1ll<<(maxDepth+1))+(1ll<<(2*(maxDepth+1))) );", "encode base (device)")
    rep(r"sonKeyY \* \(1<<\(depthD-depth\)\) \* \(1<<\(maxDepth\+1\)\) \+", "sonKeyY * (1ll<<(depthD-depth)) * (1ll<<(maxDepth+1)) +", "encode y (device)")
    rep(r"sonKeyZ \* \(1<<\(depthD-depth\)\) \* \(1<<\(2\*\(maxDepth\+1\)\)\);", "sonKeyZ * (1ll<<(depthD-depth)) * (1ll<<(2*(maxDepth+1)));", "encode z (device)")
    rep(r"nByte = 1ll \* sizeof\(int\) \* NodeArray_sz;\n(\s*)CHECK\(cudaMalloc\(\(long long \*\*\)&EncodedNodeIdxInFunction",
        r"nByte = 1ll * sizeof(long long) * NodeArray_sz;\n\1CHECK(cudaMalloc((long long **)&EncodedNodeIdxInFunction", "allocation size")
    return s


def patch_cg(src):
    s = src
    # (4) iteration counter: one store at kernel exit, printed by the host wrapper
    s = sub_once(s, 'extern "C" __global__ void gpuConjugateGradient(',
                 '__device__ int g_refCgIters;\nextern "C" __global__ void gpuConjugateGradient(', "cg iters decl")
    s = sub_once(s, "        r1 = *dot_result;\n        k++;\n    }\n}\n",
                 "        r1 = *dot_result;\n        k++;\n    }\n    if (threadIdx.x == 0 && blockIdx.x == 0) g_refCgIters = k - 1;\n}\n", "cg iters store")
    s = sub_once(s, '    printf("GPU Final, residual = %e, kernel execution time = %f ms\\n", sqrt(r1),\n           time);\n',
                 '    printf("GPU Final, residual = %e, kernel execution time = %f ms\\n", sqrt(r1),\n           time);\n'
                 '    { int it=0; cudaMemcpyFromSymbol(&it,g_refCgIters,sizeof(int)); printf("REF_CG rows=%d iters=%d ms=%f\\n",N,it,time);\n'
                 '      const char* d=getenv("REF_DUMP_DIR"); if(d){ char path[1024]; snprintf(path,sizeof(path),"%s/cg.txt",d); FILE* fp=fopen(path,"a"); if(fp){ fprintf(fp,"%d %d %f %e\\n",N,it,time,sqrt(r1)); fclose(fp);} } }\n',
                 "cg iters print")
    return s


def build_tables():
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", "-O1", "-std=c++14", "-w", "-I/usr/local/cuda/include", "-I" + REF,
           os.path.join(HERE, "ref_tables_harness.cpp"), "-o", os.path.join(OUT, "ref_tables")]
    subprocess.check_call(cmd)


def build_poisson(depth, keep=False):
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix=f"prb_ref_d{depth}_")
    try:
        for f in os.listdir(REF):
            p = os.path.join(REF, f)
            if os.path.isfile(p) and (f.endswith((".cu", ".cuh", ".inl", ".h"))):
                shutil.copy(p, os.path.join(tmp, f))
        mp = os.path.join(tmp, "main.cu")
        with open(mp) as fh:
            src = fh.read()
        with open(mp, "w") as fh:
            src = patch_main(src)
            if depth >= 10:
                src = patch_widen(src)
            fh.write(src)
        cp = os.path.join(tmp, "CG_CUDA.cuh")
        with open(cp) as fh:
            src = fh.read()
        with open(cp, "w") as fh:
            fh.write(patch_cg(src))
        out = os.path.join(OUT, f"ref_poisson_d{depth}" + ("_widen" if depth >= 10 else ""))
        cmd = ["nvcc", "-arch=sm_100", "-std=c++17", "-rdc=true", "-w", "-O2", f"-DREF_DEPTH={depth}",
               "main.cu", "CmdLineParser.cu", "Geometry.cu", "plyfile.cu", "Factor.cu", "-o", out]
        subprocess.check_call(cmd, cwd=tmp)
        return out
    finally:
        if not keep:
            shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depths", type=int, nargs="*", default=[8, 9])
    ap.add_argument("--tables-only", action="store_true")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        print("reference sources not present; nothing to build (prebuilt oracle/_ref is used)")
        return 0
    build_tables()
    if not a.tables_only:
        for d in a.depths:
            if not (5 <= d <= 10):
                raise SystemExit("the reference is only valid for 5 <= maxDepth <= 9 (10 with the 'ref+widen' index-type patch)")
            print("built", build_poisson(d, a.keep))
    return 0


if __name__ == "__main__":
    sys.exit(main())

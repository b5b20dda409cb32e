// tests/tools/prox_host.cu -- TEST/DEBUG TOOL, NOT PRODUCT CODE.
// Recompiles the device prox source (admm-elastic_b200/csrc/prox.cuh) for the host so that parity
// problems (basins of the Newton iteration, fp32 behaviour) can be studied in a container without a
// GPU.  Nothing in the package loads this library; host libm differs from device math in the last
// bits, so it is a study aid, never the thing tested for parity.
#define ADMMB200_HOST_SHIM 1
#include "../../admm-elastic_b200/csrc/prox.cuh"
using namespace admmb200;

template <typename T, int MODEL> static void run(double mu, double lambda, double kappa, int n, const double *zin, double *zout)
{
	Material<T> m = Material<T>::make(mu, lambda, kappa);
	for (int e = 0; e < n; ++e) {
		T z[9];
		for (int k = 0; k < 9; ++k) z[k] = T(zin[9 * e + k]);
		prox_tet<T, MODEL>(m, z);
		for (int k = 0; k < 9; ++k) zout[9 * e + k] = double(z[k]);
	}
}
template <typename T> static int dispatch(int model, double mu, double lambda, double kappa, int n, const double *zin, double *zout)
{
	switch (model) {
	case 0: run<T, 0>(mu, lambda, kappa, n, zin, zout); break;
	case 1: run<T, 1>(mu, lambda, kappa, n, zin, zout); break;
	case 2: run<T, 2>(mu, lambda, kappa, n, zin, zout); break;
	case 3: run<T, 3>(mu, lambda, kappa, n, zin, zout); break;
	case 4: run<T, 4>(mu, lambda, kappa, n, zin, zout); break;
	case 5: run<T, 5>(mu, lambda, kappa, n, zin, zout); break;
	default: return 1;
	}
	return 0;
}
extern "C" int shim_prox_tets(int model, double mu, double lambda, double kappa, int precision, int n, const double *zin, double *zout)
{
	return precision ? dispatch<double>(model, mu, lambda, kappa, n, zin, zout) : dispatch<float>(model, mu, lambda, kappa, n, zin, zout);
}
extern "C" int shim_prox_tris(double lmin, double lmax, int precision, int n, const double *zin, double *zout)
{
	for (int e = 0; e < n; ++e) {
		if (precision) { double z[6]; for (int k = 0; k < 6; ++k) z[k] = zin[6 * e + k]; prox_tri<double>(lmin, lmax, z); for (int k = 0; k < 6; ++k) zout[6 * e + k] = z[k]; }
		else { float z[6]; for (int k = 0; k < 6; ++k) z[k] = float(zin[6 * e + k]); prox_tri<float>(float(lmin), float(lmax), z); for (int k = 0; k < 6; ++k) zout[6 * e + k] = z[k]; }
	}
	return 0;
}

// warm-started SVD (quaternion of V carried from call to call, as tet_local_kernel does): qio is [n][4], in/out
template <typename T, int MODEL> static void run_warm(double mu, double lambda, double kappa, int n, const double *zin, double *zout, double *qio)
{
	Material<T> m = Material<T>::make(mu, lambda, kappa);
	for (int e = 0; e < n; ++e) {
		T z[9], q[4];
		for (int k = 0; k < 9; ++k) z[k] = T(zin[9 * e + k]);
		for (int k = 0; k < 4; ++k) q[k] = T(qio[4 * e + k]);
		prox_tet_mode<T, MODEL, PROX_INLINE>(m, z, q);
		for (int k = 0; k < 9; ++k) zout[9 * e + k] = double(z[k]);
		for (int k = 0; k < 4; ++k) qio[4 * e + k] = double(q[k]);
	}
}
extern "C" int shim_prox_tets_warm(int model, double mu, double lambda, double kappa, int precision, int n, const double *zin, double *zout, double *qio)
{
	if (model < 0 || model > 2) return 1;
	if (precision) { if (model == 0) run_warm<double, 0>(mu, lambda, kappa, n, zin, zout, qio); else if (model == 1) run_warm<double, 1>(mu, lambda, kappa, n, zin, zout, qio); else run_warm<double, 2>(mu, lambda, kappa, n, zin, zout, qio); }
	else { if (model == 0) run_warm<float, 0>(mu, lambda, kappa, n, zin, zout, qio); else if (model == 1) run_warm<float, 1>(mu, lambda, kappa, n, zin, zout, qio); else run_warm<float, 2>(mu, lambda, kappa, n, zin, zout, qio); }
	return 0;
}

/* Globals that the reference's main() owns (fscl.c:33-34) and its host objects import
 * (sm-spline.c:15); defined here for libfscl_hostref.so, which carries no main(). */
int spline_pts = 200;     /* --splines default, fscl.c:171 */
